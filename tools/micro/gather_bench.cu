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
// layout 2 ("quad-packed fp16"): q[b][g][y0+1][x0+1][octet] = 64-byte record {4 corners x 8 channels, fp16}, zero outside
#include <cuda_fp16.h>
__global__ void gather_quad_kernel(const float* __restrict__ q, const float* __restrict__ off, const float* __restrict__ msk,
                                   float* __restrict__ out, int B, int C, int H, int W, int DG) {
    const int cg = C / DG, P = H * W, pieces = C / 8, oct = cg / 8;
    const long long total = (long long)B * P * pieces;
    float acc = 0.f;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
        const int piece = t % pieces, p = (t / pieces) % P, b = t / ((long long)pieces * P);
        const int g = piece / oct, o = piece - g * oct;
        const int oy = p / W, ox = p % W;
#pragma unroll 3
        for (int tap = 0; tap < 9; ++tap) {
            const size_t ob = ((size_t)(b * DG + g) * 18 + 2 * tap) * P + p;
            const float y = oy - 1 + tap / 3 + __ldg(off + ob), xx = ox - 1 + tap % 3 + __ldg(off + ob + P);
            const float mk = __ldg(msk + ((size_t)(b * DG + g) * 9 + tap) * P + p);
            if (!(y > -1.f && xx > -1.f && y < H && xx < W)) continue;
            const int y0 = (int)floorf(y), x0 = (int)floorf(xx);
            const float ly = y - y0, lx = xx - x0;
            const size_t rec = ((((size_t)(b * DG + g) * (H + 1) + y0 + 1) * (W + 1) + x0 + 1) * oct + o) * 16;
            const F8 lo = ldg8(q + rec), hi = ldg8(q + rec + 8);
            const float w0 = (1 - ly) * (1 - lx) * mk, w1 = (1 - ly) * lx * mk, w2 = ly * (1 - lx) * mk, w3 = ly * lx * mk;
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&lo.v[e]));
                const float2 bb = __half22float2(*reinterpret_cast<const __half2*>(&lo.v[4 + e]));
                const float2 c = __half22float2(*reinterpret_cast<const __half2*>(&hi.v[e]));
                const float2 d = __half22float2(*reinterpret_cast<const __half2*>(&hi.v[4 + e]));
                acc += w0 * a.x + w1 * bb.x + w2 * c.x + w3 * d.x;
                acc += w0 * a.y + w1 * bb.y + w2 * c.y + w3 * d.y;
            }
        }
    }
    if (acc == 1234.5678f) out[0] = acc;
}
// quad-packed, lane pairs: each load instruction covers 16 whole 64-byte records (adjacent lanes read the two
// 32-byte halves of one record), partial blends are exchanged with shuffles
__global__ void gather_quad_pair_kernel(const float* __restrict__ q, const float* __restrict__ off, const float* __restrict__ msk,
                                        float* __restrict__ out, int B, int C, int H, int W, int DG) {
    const int cg = C / DG, P = H * W, pieces = C / 8, oct = cg / 8;
    const long long total = (long long)B * P * pieces;     // multiple of 32 here
    const int lane = threadIdx.x & 31, half = lane & 1;
    float acc = 0.f;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
        const int piece = t % pieces, p = (t / pieces) % P, b = t / ((long long)pieces * P);
        const int g = piece / oct, o = piece - g * oct;
        const int oy = p / W, ox = p % W;
#pragma unroll 3
        for (int tap = 0; tap < 9; ++tap) {
            const size_t ob = ((size_t)(b * DG + g) * 18 + 2 * tap) * P + p;
            const float y = oy - 1 + tap / 3 + __ldg(off + ob), xx = ox - 1 + tap % 3 + __ldg(off + ob + P);
            float mk = __ldg(msk + ((size_t)(b * DG + g) * 9 + tap) * P + p);
            const bool in = (y > -1.f && xx > -1.f && y < H && xx < W);
            const int y0 = in ? (int)floorf(y) : -1, x0 = in ? (int)floorf(xx) : -1;
            if (!in) mk = 0.f;
            const float ly = y - y0, lx = xx - x0;
            const unsigned rec = (unsigned)((((size_t)(b * DG + g) * (H + 1) + y0 + 1) * (W + 1) + x0 + 1) * oct + o) * 16u;
            const float wlo0 = (1 - ly) * (1 - lx) * mk, wlo1 = (1 - ly) * lx * mk, whi0 = ly * (1 - lx) * mk, whi1 = ly * lx * mk;
            const unsigned rec_e = __shfl_sync(0xffffffffu, rec, lane & ~1), rec_o = __shfl_sync(0xffffffffu, rec, lane | 1);
            const F8 ve = ldg8(q + rec_e + half * 8), vo = ldg8(q + rec_o + half * 8);
            // weights of my corner pair for the even and the odd sample
            const float ma = half ? whi0 : wlo0, mb = half ? whi1 : wlo1;       // my sample, my half's corners
            const float xa = half ? wlo0 : whi0, xb = half ? wlo1 : whi1;       // my sample, the partner's corners
            const float pa = __shfl_xor_sync(0xffffffffu, xa, 1), pb = __shfl_xor_sync(0xffffffffu, xb, 1);  // partner's sample, my corners
            const float wea = half ? pa : ma, web = half ? pb : mb, woa = half ? ma : pa, wob = half ? mb : pb;
            float se[8], so[8];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&ve.v[e]));
                const float2 bb = __half22float2(*reinterpret_cast<const __half2*>(&ve.v[4 + e]));
                const float2 c = __half22float2(*reinterpret_cast<const __half2*>(&vo.v[e]));
                const float2 d = __half22float2(*reinterpret_cast<const __half2*>(&vo.v[4 + e]));
                se[2 * e] = wea * a.x + web * bb.x; se[2 * e + 1] = wea * a.y + web * bb.y;
                so[2 * e] = woa * c.x + wob * d.x;  so[2 * e + 1] = woa * c.y + wob * d.y;
            }
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                const float give = half ? se[e] : so[e], keep = half ? so[e] : se[e];
                acc += keep + __shfl_xor_sync(0xffffffffu, give, 1);
            }
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
        const size_t nq = (size_t)B * DG * (H + 1) * (W + 1) * (C / DG / 8) * 16;   // floats (64-byte records)
        float* q; cudaMalloc(&q, nq * 4); cudaMemset(q, 0, nq * 4);
        cudaMalloc(&x, nx * 4); cudaMalloc(&off, no * 4); cudaMalloc(&msk, nm * 4); cudaMalloc(&out, 64);
        std::vector<float> h(no);
        srand(1);
        for (size_t i = 0; i < no; ++i) { float u = 0; for (int k = 0; k < 4; ++k) u += rand() / (float)RAND_MAX - 0.5f; h[i] = u * 5.2f; }  // ~N(0, 3^2)
        cudaMemcpy(off, h.data(), no * 4, cudaMemcpyHostToDevice);
        cudaMemset(x, 0, nx * 4); cudaMemset(msk, 0, nm * 4);
        for (int variant = 4; variant < 6; ++variant) {
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
                    if (variant == 5) gather_quad_pair_kernel<<<grid, threads>>>(q, off, msk, out, B, C, H, W, DG);
                    if (variant == 4) gather_quad_kernel<<<grid, threads>>>(q, off, msk, out, B, C, H, W, DG);
                    cudaEventRecord(e1); cudaEventSynchronize(e1);
                    float ms; cudaEventElapsedTime(&ms, e0, e1); if (rep > 0 && ms < best) best = ms;
                }
                printf("C=%3d hw=%3d layout=%s mapping=%s offsets=%s : %.3f ms  (%s)\n", C, H, variant == 5 ? "quad16 paired" : variant == 4 ? "quad16" : variant / 2 ? "gmajor" : "nhwc  ",
                       variant % 2 ? "lane=position" : "lane=piece   ", smooth ? "zero  " : "random", best, cudaGetErrorString(cudaGetLastError()));
            }
        }
        cudaFree(q); cudaFree(x); cudaFree(off); cudaFree(msk); cudaFree(out);
    }
    return 0;
}
