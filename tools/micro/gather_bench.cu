// Micro-benchmark: how fast can the DCN bilinear gather alone run on B200 (no MMA, no smem staging)?
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o gather_bench gather_bench.cu && ./gather_bench
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cmath>
struct __align__(32) F8 { float v[8]; };
__device__ __forceinline__ F8 ldg8(const float* p) {
    F8 r;
    asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
        : "=f"(r.v[0]), "=f"(r.v[1]), "=f"(r.v[2]), "=f"(r.v[3]), "=f"(r.v[4]), "=f"(r.v[5]), "=f"(r.v[6]), "=f"(r.v[7]) : "l"(p));
    return r;
}
// layout 0: NHWC  x[b][y][x][C];  layout 1: group-major x[b][g][y][x][cg]
// mapping 0: lane -> (piece-major: 4 pieces x 8 positions) ; mapping 1: lane -> 32 consecutive positions, fixed piece
template <int LAYOUT, int MAPPING>
__global__ void gather_kernel(const float* __restrict__ x, const float* __restrict__ off, const float* __restrict__ msk,
                              float* __restrict__ out, int B, int C, int H, int W, int DG) {
    const int cg = C / DG, P = H * W, pieces = C / 8;
    const long long total = (long long)B * P * pieces;
    float acc = 0.f;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
        int piece, p, b;
        if (MAPPING == 0) { piece = t % pieces; p = (t / pieces) % P; b = t / ((long long)pieces * P); }
        else { p = t % P; piece = (t / P) % pieces; b = t / ((long long)pieces * P); }
        const int c0 = piece * 8, g = c0 / cg;
        const int oy = p / W, ox = p % W;
#pragma unroll 3
        for (int tap = 0; tap < 9; ++tap) {
            const size_t ob = ((size_t)(b * DG + g) * 18 + 2 * tap) * P + p;
            const float y = oy - 1 + tap / 3 + __ldg(off + ob), xx = ox - 1 + tap % 3 + __ldg(off + ob + P);
            const float mk = __ldg(msk + ((size_t)(b * DG + g) * 9 + tap) * P + p);
            if (!(y > -1.f && xx > -1.f && y < H && xx < W)) continue;
            const int y0 = max(0, min(H - 2, (int)floorf(y))), x0 = max(0, min(W - 2, (int)floorf(xx)));
            const float ly = y - y0, lx = xx - x0;
            const float* base; int xs, ys;
            if (LAYOUT == 0) { base = x + (((size_t)b * H + y0) * W + x0) * C + c0; xs = C; ys = W * C; }
            else { base = x + ((((size_t)b * DG + g) * H + y0) * W + x0) * cg + (c0 - g * cg); xs = cg; ys = W * cg; }
            const F8 a = ldg8(base), bb = ldg8(base + xs), c = ldg8(base + ys), d = ldg8(base + ys + xs);
            const float w0 = (1 - ly) * (1 - lx) * mk, w1 = (1 - ly) * lx * mk, w2 = ly * (1 - lx) * mk, w3 = ly * lx * mk;
#pragma unroll
            for (int e = 0; e < 8; ++e) acc += w0 * a.v[e] + w1 * bb.v[e] + w2 * c.v[e] + w3 * d.v[e];
        }
    }
    if (acc == 1234.5678f) out[0] = acc;
}
int main(int argc, char** argv) {
    const int bps = argc > 1 ? atoi(argv[1]) : 8, threads = argc > 2 ? atoi(argv[2]) : 256;
    const int B = 80;
    const int cfgs[3][2] = {{256, 40}, {128, 80}, {64, 160}};
    for (int ci = 0; ci < 3; ++ci) {
        const int C = cfgs[ci][0], H = cfgs[ci][1], W = H, DG = 8, P = H * W;
        const size_t nx = (size_t)B * C * P, no = (size_t)B * DG * 18 * P, nm = no / 2;
        float *x, *off, *msk, *out;
        cudaMalloc(&x, nx * 4); cudaMalloc(&off, no * 4); cudaMalloc(&msk, nm * 4); cudaMalloc(&out, 64);
        std::vector<float> h(no);
        srand(1);
        for (size_t i = 0; i < no; ++i) { float u = 0; for (int k = 0; k < 4; ++k) u += rand() / (float)RAND_MAX - 0.5f; h[i] = u * 5.2f; }  // ~N(0, 3^2)
        cudaMemcpy(off, h.data(), no * 4, cudaMemcpyHostToDevice);
        cudaMemset(x, 0, nx * 4); cudaMemset(msk, 0, nm * 4);
        for (int variant = 0; variant < 1; ++variant) {
            for (int smooth = 0; smooth < 2; ++smooth) {
                if (smooth) cudaMemset(off, 0, no * 4); else cudaMemcpy(off, h.data(), no * 4, cudaMemcpyHostToDevice);
                cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
                float best = 1e9;
                for (int rep = 0; rep < 4; ++rep) {
                    cudaEventRecord(e0);
                    const int grid = 148 * bps;
                    if (variant == 0) gather_kernel<0, 0><<<grid, threads>>>(x, off, msk, out, B, C, H, W, DG);
                    if (variant == 1) gather_kernel<0, 1><<<grid, threads>>>(x, off, msk, out, B, C, H, W, DG);
                    if (variant == 2) gather_kernel<1, 0><<<grid, threads>>>(x, off, msk, out, B, C, H, W, DG);
                    if (variant == 3) gather_kernel<1, 1><<<grid, threads>>>(x, off, msk, out, B, C, H, W, DG);
                    cudaEventRecord(e1); cudaEventSynchronize(e1);
                    float ms; cudaEventElapsedTime(&ms, e0, e1); if (rep > 0 && ms < best) best = ms;
                }
                printf("C=%3d hw=%3d layout=%s mapping=%s offsets=%s : %.3f ms  (%s)\n", C, H, variant / 2 ? "gmajor" : "nhwc  ",
                       variant % 2 ? "lane=position" : "lane=piece   ", smooth ? "zero  " : "random", best, cudaGetErrorString(cudaGetLastError()));
            }
        }
        cudaFree(x); cudaFree(off); cudaFree(msk); cudaFree(out);
    }
    return 0;
}
