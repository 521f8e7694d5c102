// Micro-benchmark: DCN corner fetch through the TMA engine (cp.async.bulk.tensor ... tile::gather4) instead of the LSU.
// One gather4 = the four bilinear corners (4 pixel rows of an [B*H*W, C] NHWC matrix) x one deform group's channels.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tma_gather_bench tma_gather_bench.cu && ./tma_gather_bench
// args: warps-per-CTA  ring-depth  box-rows(1|4)  ctas-per-sm
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, int n) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(n)); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* b, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* b, uint32_t parity) {
    uint32_t ok;
    asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                 : "=r"(ok) : "r"(smem_u32(b)), "r"(parity) : "memory");
    return ok;
}
__device__ __forceinline__ void tma_gather4(void* dst, const CUtensorMap* map, uint64_t* bar, int col, int r0, int r1, int r2, int r3) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile::gather4.mbarrier::complete_tx::bytes"
                 " [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
                 :: "r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(col), "r"(r0), "r"(r1), "r"(r2), "r"(r3) : "memory");
}
__device__ __forceinline__ uint32_t hash32(uint32_t x) {
    x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16;
    return x;
}
// op id -> (top-left pixel row index, channel column)
__device__ __forceinline__ void op_coords(long long op, int H, int W, int DG, int cg, int spread, int& row, int& col) {
    const int g = (int)(op % DG);
    long long t = op / DG;
    const int tap = (int)(t % 9); t /= 9;
    const int p = (int)(t % (H * W)); const int b = (int)(t / (H * W));
    const int oy = p / W, ox = p % W;
    const uint32_t h = hash32((uint32_t)op * 2654435761u + 12345u);
    int y0 = oy - 1 + tap / 3 + (int)(h % (2 * spread + 1)) - spread;
    int x0 = ox - 1 + tap % 3 + (int)((h >> 12) % (2 * spread + 1)) - spread;
    y0 = max(0, min(H - 2, y0)); x0 = max(0, min(W - 2, x0));
    row = (b * H + y0) * W + x0;
    col = g * cg;
}

template <int CG>   // floats per gathered row piece
__global__ void __launch_bounds__(1024) tma_gather_kernel(const __grid_constant__ CUtensorMap map, float* __restrict__ out,
                                                          long long n_ops, int H, int W, int DG, int spread, int depth) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    constexpr int OPB = 4 * CG * 4;                  // bytes per gather4
    uint8_t* ring = smem + (size_t)warp * depth * 32 * OPB;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)nw * depth * 32 * OPB) + warp * depth;
    if (lane == 0) for (int i = 0; i < depth; ++i) mbar_init(&bars[i], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncthreads();
    const long long gw = (long long)blockIdx.x * nw + warp, tw = (long long)gridDim.x * nw;
    const long long iters = (n_ops / 32 + tw - 1 - gw) / tw;   // warp-iterations of 32 ops
    float acc = 0.f;
    auto issue = [&](long long it) {
        const int slot = (int)(it % depth);
        const long long op = (gw + it * tw) * 32 + lane;
        int row, col;
        op_coords(op, H, W, DG, CG, spread, row, col);
        if (lane == 0) mbar_expect_tx(&bars[slot], 32 * OPB);
        __syncwarp();
        tma_gather4(ring + ((size_t)slot * 32 + lane) * OPB, &map, &bars[slot], col, row, row + 1, row + W, row + W + 1);
    };
    for (long long it = 0; it < depth - 1 && it < iters; ++it) issue(it);
    for (long long it = 0; it < iters; ++it) {
        if (it + depth - 1 < iters) issue(it + depth - 1);
        const int slot = (int)(it % depth);
        while (!mbar_try_wait(&bars[slot], (uint32_t)((it / depth) & 1))) {}
        const float4* s = reinterpret_cast<const float4*>(ring + ((size_t)slot * 32 + lane) * OPB);
#pragma unroll
        for (int e = 0; e < OPB / 16; ++e) { const float4 v = s[e]; acc += v.x + v.y + v.z + v.w; }
        __syncwarp();
    }
    atomicAdd(out, acc);
}

template <int CG>
__global__ void lsu_gather_kernel(const float* __restrict__ x, float* __restrict__ out, long long n_ops, int H, int W, int DG,
                                  int C, int spread) {
    float acc = 0.f;
    for (long long op = (long long)blockIdx.x * blockDim.x + threadIdx.x; op < n_ops; op += (long long)gridDim.x * blockDim.x) {
        int row, col;
        op_coords(op, H, W, DG, CG, spread, row, col);
        const int rows[4] = {row, row + 1, row + W, row + W + 1};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const float4* p = reinterpret_cast<const float4*>(x + (size_t)rows[k] * C + col);
#pragma unroll
            for (int e = 0; e < CG / 4; ++e) { const float4 v = __ldg(p + e); acc += v.x + v.y + v.z + v.w; }
        }
    }
    atomicAdd(out, acc);
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

template <int CG>
void run(int C, int HW, int nwarps, int depth, int box_rows, int cps, EncodeFn enc) {
    const int B = 80, H = HW, W = HW, DG = 8;
    const size_t rows = (size_t)B * H * W, nx = rows * C;
    float *x, *out;
    CK(cudaMalloc(&x, nx * 4)); CK(cudaMalloc(&out, 8));
    std::vector<float> h(nx);
    for (size_t i = 0; i < nx; ++i) h[i] = (float)((i * 2654435761u >> 20) & 255) / 256.f;
    CK(cudaMemcpy(x, h.data(), nx * 4, cudaMemcpyHostToDevice));
    CUtensorMap map;
    cuuint64_t dims[2] = {(cuuint64_t)C, (cuuint64_t)rows}, strides[1] = {(cuuint64_t)C * 4};
    cuuint32_t box[2] = {(cuuint32_t)CG, (cuuint32_t)box_rows}, es[2] = {1, 1};
    CUresult r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, x, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("C=%d encode failed %d (box rows %d)\n", C, (int)r, box_rows); return; }
    const long long n_ops = (long long)B * H * W * 9 * DG;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float ms;
    for (int spread : {8, 0}) {
        // reference sum through the LSU
        CK(cudaMemset(out, 0, 4));
        lsu_gather_kernel<CG><<<148 * 8, 256>>>(x, out, n_ops, H, W, DG, C, spread);
        CK(cudaDeviceSynchronize());
        cudaEventRecord(e0);
        for (int i = 0; i < 5; ++i) lsu_gather_kernel<CG><<<148 * 8, 256>>>(x, out, n_ops, H, W, DG, C, spread);
        cudaEventRecord(e1); CK(cudaDeviceSynchronize()); cudaEventElapsedTime(&ms, e0, e1);
        const float lsu_ms = ms / 5;
        CK(cudaMemset(out, 0, 4));
        lsu_gather_kernel<CG><<<148 * 8, 256>>>(x, out, n_ops, H, W, DG, C, spread);
        float ref; CK(cudaMemcpy(&ref, out, 4, cudaMemcpyDeviceToHost));
        const size_t smem = (size_t)nwarps * depth * 32 * 16 * CG + nwarps * depth * 8 + 64;
        CK(cudaFuncSetAttribute(tma_gather_kernel<CG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        CK(cudaMemset(out, 0, 4));
        tma_gather_kernel<CG><<<148 * cps, nwarps * 32, smem>>>(map, out, n_ops, H, W, DG, spread, depth);
        CK(cudaDeviceSynchronize());
        float got; CK(cudaMemcpy(&got, out, 4, cudaMemcpyDeviceToHost));
        cudaEventRecord(e0);
        for (int i = 0; i < 5; ++i) tma_gather_kernel<CG><<<148 * cps, nwarps * 32, smem>>>(map, out, n_ops, H, W, DG, spread, depth);
        cudaEventRecord(e1); CK(cudaDeviceSynchronize()); cudaEventElapsedTime(&ms, e0, e1);
        printf("C=%3d hw=%3d piece=%3dB spread=%d  warps=%d depth=%d ctas/sm=%d smem=%zuKB : tma gather4 %.3f ms (%.2f ops/clk/SM @1.9GHz)  lsu %.3f ms  sum tma %.6g lsu %.6g %s\n",
               C, HW, CG * 4, spread, nwarps, depth, cps, smem / 1024, ms / 5, n_ops / (ms / 5 * 1e-3) / 148 / 1.9e9, lsu_ms, got, ref,
               fabsf(got - ref) <= 1e-3f * fabsf(ref) ? "OK" : "MISMATCH");
    }
    cudaFree(x); cudaFree(out);
}

int main(int argc, char** argv) {
    const int nwarps = argc > 1 ? atoi(argv[1]) : 8, depth = argc > 2 ? atoi(argv[2]) : 4;
    const int box_rows = argc > 3 ? atoi(argv[3]) : 1, cps = argc > 4 ? atoi(argv[4]) : 1;
    EncodeFn enc = nullptr;
    cudaDriverEntryPointQueryResult q;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void**)&enc, cudaEnableDefault, &q));
    run<8>(64, 160, nwarps, depth, box_rows, cps, enc);
    run<16>(128, 80, nwarps, depth, box_rows, cps, enc);
    run<32>(256, 40, nwarps, depth, box_rows, cps, enc);
    return 0;
}
