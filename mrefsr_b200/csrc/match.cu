// mrefsr_b200/csrc/match.cu -- correspondence matcher (feature_match_index), batched over (image, ref) pairs.
//
// Replaces basicsr/archs/ref_map_util.py:26-86 (and the per-pixel F.normalize of
// basicsr/archs/corres_generation_arch.py:57-59 when normalize_pixels is set).
//
// Pipeline (all on one stream):
//   1. match_prep_kernel      NCHW fp32 -> pixel-major [HW, C] copy (fp32, or split bf16 hi/lo), optional
//                             per-pixel L2 normalisation, per-pixel sum of squares.
//   2. match_norms_kernel     3x3 (ps x ps) box sums of the pixel sums -> reference-patch 1/(norm+1e-5) as a
//                             per-column (scale, bias) pair (bias = -inf masks columns that are not a
//                             valid patch origin), input-patch (norm+1e-5) per row.
//   3. match_tc_kernel        the correlation as a TMA-fed tcgen05/TMEM GEMM with implicit im2col by row
//      (or match_simt_kernel)  shift, arg-max fused into the epilogue: the similarity volume never leaves
//                             the SM.  Partial (max, argmax) per N tile are merged with one 64-bit
//                             atomicMax per row on a packed key (ordered value << 32 | ~column).
//   4. match_finalize_kernel  key -> (int64 index on the reference patch grid, fp32 value / input norm).
//
// Implicit im2col: with pixel-major features F[HW, C], the K = 9*C patch vector of the patch whose origin is
// pixel-linear index p is the concatenation over taps (i, j) of row p + i*w + j.  So the A tile of tap (i, j)
// for rows [m0, m0+128) is simply rows [m0 + i*w + j, ...) of F: a plain 2-D TMA box, no unfold copy.
// Rows whose origin is not a valid patch origin (x > w-3 or y > h-3) are computed and discarded (10 % at 40x40).
#include "common.cuh"
#include "../../include/mrefsr_b200.h"

namespace mrefsr {

// 16-bit operand type of the tensor-core path.  The hi / lo halves of the split (x = hi + lo) are bf16 by default;
// -DMREFSR_MATCH_SPLIT_FP16 builds them as fp16 instead: same MMA cost (kind::f16 takes either), and in the CPU
// emulation (tools/match_two_sweep_sim.py) a 3.6x smaller worst-case similarity error (1.7e-6 vs 6.1e-6), because
// the unit-normalised features need none of bf16's exponent range.  Not yet run on hardware -> not the default.
#ifdef MREFSR_MATCH_SPLIT_FP16
typedef __half bf16;                 // (the name stays: "the 16-bit half of the split")
typedef __half2 bf16x2;
#define MREFSR_TO16(v) __float2half_rn(v)
#define MREFSR_FROM16(h) __half2float(h)
#define MREFSR_TMAP_16 CU_TENSOR_MAP_DATA_TYPE_FLOAT16
#define MREFSR_IDESC_16 0
#else
typedef __nv_bfloat16 bf16;
typedef __nv_bfloat162 bf16x2;
#define MREFSR_TO16(v) __float2bfloat16_rn(v)
#define MREFSR_FROM16(h) __bfloat162float(h)
#define MREFSR_TMAP_16 CU_TENSOR_MAP_DATA_TYPE_BFLOAT16
#define MREFSR_IDESC_16 1
#endif

// =====================================================================================================
// 1. prep
// =====================================================================================================
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// grid (ceil(HW/32), n_img), block 256, dynamic smem C*33 floats.
// Phase 1: the block's 32 pixels x C channels through a padded shared-memory tile, sixteen independent 128-byte row
// loads in flight per warp (the pass is HBM-bound; a dependent load-store loop leaves the memory system idle).
// Phase 2: one warp per pixel, the pixel's channels (pairs 2*lane + 64*k) in registers from one tile read: sum of
// squares, optional normalisation, hi / lo split and the stores -- no second trip through shared memory.
template <int CPL>   // channel pairs per lane = C / 64 (0: generic loop over the tile for other channel counts)
__global__ void __launch_bounds__(256)
match_prep_kernel(const float* __restrict__ src, int C, int HW, int normalize, float* __restrict__ dst_f32,
                  bf16* __restrict__ dst_hi, bf16* __restrict__ dst_lo, float* __restrict__ pix_sumsq) {
    extern __shared__ float tile[];  // [C][33]
    const int img = blockIdx.y, p0 = blockIdx.x * 32;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const float* s = src + (size_t)img * C * HW;
    const int p = p0 + lane;
    constexpr int U = 16;
#pragma unroll 1
    for (int c0 = warp; c0 < C; c0 += nwarps * U) {
        float v[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int c = c0 + u * nwarps;
            v[u] = (p < HW && c < C) ? __ldg(s + (size_t)c * HW + p) : 0.f;
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int c = c0 + u * nwarps;
            if (c < C) tile[c * 33 + lane] = v[u];
        }
    }
    __syncthreads();
    for (int pp = warp; pp < 32; pp += nwarps) {
        const int q = p0 + pp;
        if (CPL > 0) {
            float2 v[CPL > 0 ? CPL : 1];
            float ss = 0.f;
#pragma unroll
            for (int k = 0; k < CPL; ++k) {
                const int c = 2 * lane + 64 * k;
                v[k].x = tile[c * 33 + pp];
                v[k].y = tile[(c + 1) * 33 + pp];
                ss = fmaf(v[k].x, v[k].x, ss);
                ss = fmaf(v[k].y, v[k].y, ss);
            }
            ss = warp_sum(ss);
            if (normalize) {
                // F.normalize(dim=0): x / max(||x||_2, 1e-12)  (corres_generation_arch.py:57-59)
                const float denom = fmaxf(sqrtf(ss), 1e-12f);
                float ss2 = 0.f;
#pragma unroll
                for (int k = 0; k < CPL; ++k) {
                    v[k].x = v[k].x / denom;
                    v[k].y = v[k].y / denom;
                    ss2 = fmaf(v[k].x, v[k].x, ss2);
                    ss2 = fmaf(v[k].y, v[k].y, ss2);
                }
                ss = warp_sum(ss2);
            }
            if (q >= HW) continue;
            if (lane == 0) pix_sumsq[(size_t)img * HW + q] = ss;
            const size_t row = ((size_t)img * HW + q) * C;
#pragma unroll
            for (int k = 0; k < CPL; ++k) {
                const int c = 2 * lane + 64 * k;
                if (dst_f32) *reinterpret_cast<float2*>(dst_f32 + row + c) = v[k];
                if (dst_hi) {
                    const bf16 h0 = MREFSR_TO16(v[k].x), h1 = MREFSR_TO16(v[k].y);
                    bf16x2 hh;
                    hh.x = h0;
                    hh.y = h1;
                    *reinterpret_cast<bf16x2*>(dst_hi + row + c) = hh;
                    if (dst_lo) {
                        bf16x2 ll;
                        ll.x = MREFSR_TO16(v[k].x - MREFSR_FROM16(h0));
                        ll.y = MREFSR_TO16(v[k].y - MREFSR_FROM16(h1));
                        *reinterpret_cast<bf16x2*>(dst_lo + row + c) = ll;
                    }
                }
            }
        } else {
            float ss = 0.f;
            for (int c = lane; c < C; c += 32) {
                float v = tile[c * 33 + pp];
                ss += v * v;
            }
            ss = warp_sum(ss);
            if (normalize) {
                const float denom = fmaxf(sqrtf(ss), 1e-12f);
                float ss2 = 0.f;
                for (int c = lane; c < C; c += 32) {
                    float v = tile[c * 33 + pp] / denom;
                    tile[c * 33 + pp] = v;
                    ss2 += v * v;
                }
                ss = warp_sum(ss2);
            }
            if (q < HW) {
                if (lane == 0) pix_sumsq[(size_t)img * HW + q] = ss;
                const size_t row = ((size_t)img * HW + q) * C;
                if (dst_f32) {
                    for (int c = lane; c < C; c += 32) dst_f32[row + c] = tile[c * 33 + pp];
                }
                if (dst_hi) {
                    for (int c = lane * 2; c < C; c += 64) {
                        const float v0 = tile[c * 33 + pp], v1 = tile[(c + 1) * 33 + pp];
                        const bf16 h0 = MREFSR_TO16(v0), h1 = MREFSR_TO16(v1);
                        bf16x2 hh;
                        hh.x = h0;
                        hh.y = h1;
                        *reinterpret_cast<bf16x2*>(dst_hi + row + c) = hh;
                        if (dst_lo) {
                            bf16x2 ll;
                            ll.x = MREFSR_TO16(v0 - MREFSR_FROM16(h0));
                            ll.y = MREFSR_TO16(v1 - MREFSR_FROM16(h1));
                            *reinterpret_cast<bf16x2*>(dst_lo + row + c) = ll;
                        }
                    }
                }
            }
        }
    }
}

// =====================================================================================================
// 2. patch norms
// =====================================================================================================
// One thread per pixel-linear position of one image.  A position is a valid patch origin when it lies on the
// stride grid and the ps x ps window fits.  ref side: colsb[img][n] = (scale, bias); input side: rownorm.
__global__ void match_norms_kernel(const float* __restrict__ pix_sumsq, int n_img, int h, int w, int ps, int stride,
                                   int padded, int is_ref, int use_norm, float2* __restrict__ colsb,
                                   float* __restrict__ rownorm) {
    const int img = blockIdx.y;
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= padded) return;
    const int HW = h * w;
    bool valid = n < HW;
    float ss = 0.f;
    if (valid) {
        const int y = n / w, x = n % w;
        valid = (y % stride == 0) && (x % stride == 0) && (y + ps <= h) && (x + ps <= w);
        if (valid) {
            const float* s = pix_sumsq + (size_t)img * HW;
            for (int i = 0; i < ps; ++i)
                for (int j = 0; j < ps; ++j) ss += s[(y + i) * w + x + j];
        }
    }
    const float nrm = sqrtf(ss) + 1e-5f;  // ref_map_util.py:63 / :79
    if (is_ref) {
        float2 v;
        v.x = valid ? (use_norm ? 1.f / nrm : 1.f) : 0.f;
        v.y = valid ? 0.f : -INFINITY;
        colsb[(size_t)img * padded + n] = v;
    } else {
        if (n < HW) rownorm[(size_t)img * HW + n] = use_norm ? nrm : 1.f;
    }
}

// =====================================================================================================
// 4. finalize
// =====================================================================================================
__global__ void match_finalize_kernel(const unsigned long long* __restrict__ keys, const float* __restrict__ rownorm,
                                      int n_pairs, int in_div, int n_in, int h_in, int w_in, int w_ref, int ho,
                                      int wo, int wo_ref, int s_in, int s_ref, int key_stride,
                                      long long* __restrict__ max_idx, float* __restrict__ max_val) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const int per = ho * wo;
    if (t >= n_pairs * per) return;
    const int pair = t / per, r = t % per;
    const int yo = r / wo, xo = r % wo;
    const int m = (yo * s_in) * w_in + xo * s_in;
    const unsigned long long key = keys[(size_t)pair * key_stride + m];
    const int img = (pair / in_div) % n_in;
    float val;
    long long idx;
    if (key == 0ull) {  // no finite similarity was ever recorded for this row
        val = __int_as_float(0x7fc00000);
        idx = 0;
    } else {
        val = ordered_to_float((uint32_t)(key >> 32));
        const uint32_t n = 0xFFFFFFFFu - (uint32_t)(key & 0xFFFFFFFFull);
        const int yq = n / w_ref, xq = n % w_ref;
        idx = (long long)(yq / s_ref) * wo_ref + xq / s_ref;
        val = val / rownorm[(size_t)img * (h_in * w_in) + m];  // ref_map_util.py:84 (rownorm == 1 if !norm_input)
    }
    max_idx[t] = idx;
    max_val[t] = val;
}

__device__ __forceinline__ unsigned long long pack_key(float v, uint32_t n) {
    return ((unsigned long long)float_to_ordered(v) << 32) | (unsigned long long)(0xFFFFFFFFu - n);
}

// =====================================================================================================
// 3a. exact-fp32 CUDA-core matcher (any patch size / stride): 64x64 tile, 4x4 micro-tile per thread
// =====================================================================================================
constexpr int ST = 64;   // tile edge
constexpr int SK = 16;   // channels per K step

__global__ void __launch_bounds__(256)
match_simt_kernel(const float* __restrict__ fin, const float* __restrict__ fref, const float2* __restrict__ colsb,
                  unsigned long long* __restrict__ keys, int in_div, int n_in, int C, int h_in, int w_in, int h_ref,
                  int w_ref, int ps, int s_in, int s_ref, int ho, int wo, int ho_ref, int wo_ref, int key_stride,
                  int colsb_stride) {
    __shared__ float As[SK][ST + 4];
    __shared__ float Bs[SK][ST + 4];
    __shared__ int rowbase[ST], colbase[ST];
    __shared__ unsigned long long skeys[ST];
    const int pair = blockIdx.z, img = (pair / in_div) % n_in;
    const int m0 = blockIdx.y * ST, n0 = blockIdx.x * ST;
    const int Nin = ho * wo, Nref = ho_ref * wo_ref;
    const int t = threadIdx.x;
    if (t < ST) {
        const int m = m0 + t;
        rowbase[t] = (m < Nin) ? ((m / wo) * s_in) * w_in + (m % wo) * s_in : -1;
        const int n = n0 + t;
        colbase[t] = (n < Nref) ? ((n / wo_ref) * s_ref) * w_ref + (n % wo_ref) * s_ref : -1;
        skeys[t] = 0ull;
    }
    __syncthreads();
    const float* A = fin + (size_t)img * h_in * w_in * C;
    const float* B = fref + (size_t)pair * h_ref * w_ref * C;
    const int lr = t >> 2, lk = (t & 3) * 4;       // loader: row, first of 4 channels
    const int ty = t >> 4, tx = t & 15;            // compute: 4 rows ty*4.., 4 cols tx*4..
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    const int rb = rowbase[lr], cb = colbase[lr];
    for (int tap = 0; tap < ps * ps; ++tap) {
        const int ti = tap / ps, tj = tap % ps;
        const float* ar = (rb >= 0) ? A + (size_t)(rb + ti * w_in + tj) * C : nullptr;
        const float* br = (cb >= 0) ? B + (size_t)(cb + ti * w_ref + tj) * C : nullptr;
        for (int c0 = 0; c0 < C; c0 += SK) {
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int c = c0 + lk + e;
                As[lk + e][lr] = (ar && c < C) ? __ldg(ar + c) : 0.f;
                Bs[lk + e][lr] = (br && c < C) ? __ldg(br + c) : 0.f;
            }
            __syncthreads();
#pragma unroll
            for (int k = 0; k < SK; ++k) {
                const float4 a = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
                const float4 b = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
                const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
            }
            __syncthreads();
        }
    }
    const float2* sb = colsb + (size_t)pair * colsb_stride;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        float best = -INFINITY;
        uint32_t bestn = 0;
        bool any = false;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int cbp = colbase[tx * 4 + j];
            if (cbp < 0) continue;
            const float2 s = sb[cbp];
            const float v = fmaf(acc[i][j], s.x, s.y);
            if (v > best) {
                best = v;
                bestn = (uint32_t)cbp;
                any = true;
            }
        }
        if (any) atomicMax(&skeys[ty * 4 + i], pack_key(best, bestn));
    }
    __syncthreads();
    if (t < ST && rowbase[t] >= 0 && skeys[t] != 0ull)
        atomicMax(keys + (size_t)pair * key_stride + rowbase[t], skeys[t]);
}

// =====================================================================================================
// 3b. tcgen05 / TMEM / TMA matcher
// =====================================================================================================
// Tile: 128 input rows x 256 reference columns, K step 64 channels (128-byte bf16 rows, SWIZZLE_128B).
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer (one elected lane),
// warps 2..5 = epilogue (TMEM lane quarter = warp_id % 4).  Two 256-column fp32 accumulators in TMEM so the
// epilogue of item k overlaps the MMAs of item k+1.  Persistent grid; work item = (pair, m_tile, n_tile).
//
// STRIP = 3: one pipeline stage holds the rows of a whole tap row (i, 0..2): 128+8 rows of A and 256+8 rows of
// B per slab; the three taps are three MMAs groups whose descriptors start j rows (j*128 bytes) into the
// strip.  This cuts L2->SMEM traffic 3x (the bf16x3 kernel would otherwise need ~62 B/clk/SM, above the
// ~42 B/clk/SM the L2 can feed all 148 SMs).  STRIP = 1 is the plain one-tap-per-stage layout.
// NPASS = 3: split-bf16 (hi*hi + hi*lo + lo*hi) for fp32-grade similarities; NPASS = 1: single bf16 pass.
template <int STRIP, int NPASS>
struct TcCfg {
    static constexpr int BM = 128, BN = 256, BK = 64;
    static constexpr int A_ROWS = (STRIP == 1) ? BM : BM + 8;
    static constexpr int B_ROWS = (STRIP == 1) ? BN : BN + 8;
    static constexpr int A_BYTES = A_ROWS * 128, B_BYTES = B_ROWS * 128;
    static constexpr int NSPLIT = (NPASS == 1) ? 1 : 2;  // hi only, or hi + lo
    static constexpr int STAGE_BYTES = NSPLIT * (A_BYTES + B_BYTES);
    static constexpr int STAGES = (200 * 1024) / STAGE_BYTES;
    static constexpr int GROUPS = 9 / STRIP;  // pipeline stages per channel slab
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 2 * BN * 8 /*scale,bias*/ + 256 /*barriers*/ + 1024;
};

struct TcParams {
    int n_pairs, in_div, n_in, C;
    int hw_in, w_in, hw_ref, w_ref;
    int m_tiles, n_tiles, total_items;
    int key_stride, colsb_stride;
    int base_offset_mode;
    int strip_rows, b_bytes, stage_bytes, stages;   // match_diag_strip_kernel: runtime stage geometry
};

template <int STRIP, int NPASS>
__global__ void __launch_bounds__(192, 1)
match_tc_kernel(const __grid_constant__ CUtensorMap mapA_hi, const __grid_constant__ CUtensorMap mapA_lo,
                const __grid_constant__ CUtensorMap mapB_hi, const __grid_constant__ CUtensorMap mapB_lo,
                const __grid_constant__ CUtensorMap mapB2_hi, const __grid_constant__ CUtensorMap mapB2_lo,
                const float2* __restrict__ colsb, unsigned long long* __restrict__ keys, const TcParams prm) {
    using Cfg = TcCfg<STRIP, NPASS>;
    constexpr int BM = Cfg::BM, BN = Cfg::BN, STAGES = Cfg::STAGES;
    // 1024-byte aligned dynamic shared memory (SWIZZLE_128B atoms); no pointer<->integer round trip, so the
    // compiler keeps every access in the shared address space (LDS/STS, 32-bit addressing)
    extern __shared__ __align__(1024) uint8_t smem[];
    if ((smem_u32(smem) & 1023u) != 0u) __trap();
    uint8_t* stage_base = smem;
    float2* s_sb = reinterpret_cast<float2*>(smem + STAGES * Cfg::STAGE_BYTES);  // [2][BN]
    uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(s_sb) + 2 * BN * 8);
    uint64_t* full = bars;                 // [STAGES]
    uint64_t* empty = bars + STAGES;       // [STAGES]
    uint64_t* tfull = bars + 2 * STAGES;   // [2]
    uint64_t* tempty = tfull + 2;          // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_slabs = prm.C / Cfg::BK;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&mapA_hi);
        tma_prefetch_desc(&mapB_hi);
        if (NPASS > 1) {
            tma_prefetch_desc(&mapA_lo);
            tma_prefetch_desc(&mapB_lo);
        }
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], 1);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(&tfull[a], 1);
            mbar_init(&tempty[a], 4);  // one arrival per epilogue warp
        }
        fence_mbar_init();
    }
    if (warp == 1) tmem_alloc<512>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ------------------------------------------------------------------ TMA producer
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int item = blockIdx.x; item < prm.total_items; item += gridDim.x) {
                const int nt = item % prm.n_tiles;
                const int mt = (item / prm.n_tiles) % prm.m_tiles;
                const int pair = item / (prm.n_tiles * prm.m_tiles);
                const int img = (pair / prm.in_div) % prm.n_in;
                const int m0 = mt * BM, n0 = nt * BN;
                for (int slab = 0; slab < n_slabs; ++slab) {
                    for (int g = 0; g < Cfg::GROUPS; ++g) {
                        const int ti = (STRIP == 3) ? g : g / 3, tj = (STRIP == 3) ? 0 : g % 3;
                        const int ra = m0 + ti * prm.w_in + tj, rb = n0 + ti * prm.w_ref + tj;
                        mbar_wait_backoff(&empty[stage], phase ^ 1, 64);
                        uint8_t* sa = stage_base + stage * Cfg::STAGE_BYTES;
                        uint8_t* sbm = sa + Cfg::NSPLIT * Cfg::A_BYTES;
                        mbar_expect_tx(&full[stage], Cfg::STAGE_BYTES);
                        tma_load_3d(sa, &mapA_hi, &full[stage], slab * Cfg::BK, ra, img);
                        tma_load_3d(sbm, &mapB_hi, &full[stage], slab * Cfg::BK, rb, pair);
                        if (STRIP == 3) tma_load_3d(sbm + BN * 128, &mapB2_hi, &full[stage], slab * Cfg::BK, rb + BN, pair);
                        if (NPASS > 1) {
                            tma_load_3d(sa + Cfg::A_BYTES, &mapA_lo, &full[stage], slab * Cfg::BK, ra, img);
                            tma_load_3d(sbm + Cfg::B_BYTES, &mapB_lo, &full[stage], slab * Cfg::BK, rb, pair);
                            if (STRIP == 3)
                                tma_load_3d(sbm + Cfg::B_BYTES + BN * 128, &mapB2_lo, &full[stage], slab * Cfg::BK,
                                            rb + BN, pair);
                        }
                        if (++stage == STAGES) {
                            stage = 0;
                            phase ^= 1;
                        }
                    }
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ------------------------------------------------------------------ MMA issuer
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            int it = 0;
            for (int item = blockIdx.x; item < prm.total_items; item += gridDim.x, ++it) {
                const int nt = item % prm.n_tiles;
                const int acc = it & 1;
                const uint32_t acc_phase = (it >> 1) & 1;
                int ncols = prm.hw_ref - nt * BN;
                ncols = ncols > BN ? BN : ((ncols + 31) & ~31);
                const uint32_t idesc = umma_idesc(MREFSR_IDESC_16, BM, ncols);
                mbar_wait_backoff(&tempty[acc], acc_phase ^ 1, 64);
                tc_fence_after();
                const uint32_t tacc = tmem_base + acc * BN;
                uint32_t accumulate = 0;
                for (int slab = 0; slab < n_slabs; ++slab) {
                    for (int g = 0; g < Cfg::GROUPS; ++g) {
                        mbar_wait(&full[stage], phase);
                        tc_fence_after();
                        const uint32_t sa = smem_u32(stage_base + stage * Cfg::STAGE_BYTES);
                        const uint32_t sb = sa + Cfg::NSPLIT * Cfg::A_BYTES;
#pragma unroll
                        for (int j = 0; j < STRIP; ++j) {
#pragma unroll
                            for (int pass = 0; pass < NPASS; ++pass) {
                                // pass 0: hi*hi, pass 1: hi*lo, pass 2: lo*hi
                                const uint32_t a0 = sa + ((pass == 2) ? Cfg::A_BYTES : 0) + j * 128;
                                const uint32_t b0 = sb + ((pass == 1) ? Cfg::B_BYTES : 0) + j * 128;
                                const uint32_t bo = prm.base_offset_mode ? (uint32_t)j : 0u;
#pragma unroll
                                for (int kk = 0; kk < 4; ++kk) {
                                    umma_f16(tacc, umma_desc_sw128(a0 + kk * 32, bo), umma_desc_sw128(b0 + kk * 32, bo),
                                             idesc, accumulate);
                                    accumulate = 1;
                                }
                            }
                        }
                        umma_commit(&empty[stage]);
                        if (++stage == STAGES) {
                            stage = 0;
                            phase ^= 1;
                        }
                    }
                }
                umma_commit(&tfull[acc]);
            }
        }
        __syncwarp();
    } else {
        // ------------------------------------------------------------------ epilogue (warps 2..5)
        const int q = warp & 3;                     // TMEM lane quarter this warp may access
        const int row = q * 32 + lane;
        const int et = threadIdx.x - 64;            // 0..127
        int it = 0;
        for (int item = blockIdx.x; item < prm.total_items; item += gridDim.x, ++it) {
            const int nt = item % prm.n_tiles;
            const int mt = (item / prm.n_tiles) % prm.m_tiles;
            const int pair = item / (prm.n_tiles * prm.m_tiles);
            const int acc = it & 1;
            const uint32_t acc_phase = (it >> 1) & 1;
            const int n0 = nt * BN;
            int ncols = prm.hw_ref - n0;
            ncols = ncols > BN ? BN : ((ncols + 31) & ~31);
            float2* sbuf = s_sb + acc * BN;
            const float2* gsb = colsb + (size_t)pair * prm.colsb_stride + n0;
            for (int i = et; i < ncols; i += 128) sbuf[i] = gsb[i];
            named_bar_sync(1, 128);
            mbar_wait_backoff(&tfull[acc], acc_phase, 512);
            tc_fence_after();
            float best = -INFINITY;
            int bestn = 0;
            const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + acc * BN;
            for (int c0 = 0; c0 < ncols; c0 += 32) {
                uint32_t v[32];
                tmem_ld_32x32(taddr + c0, v);
                tmem_ld_wait();
#pragma unroll
                for (int e = 0; e < 32; ++e) {
                    const float2 s = sbuf[c0 + e];
                    const float val = fmaf(__uint_as_float(v[e]), s.x, s.y);
                    if (val > best) {
                        best = val;
                        bestn = c0 + e;
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty[acc]);
            const int m = mt * BM + row;
            if (m < prm.hw_in && best > -INFINITY)
                atomicMax(keys + (size_t)pair * prm.key_stride + m, pack_key(best, (uint32_t)(n0 + bestn)));
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc<512>(tmem_base);
    }
}

// =====================================================================================================
// 3c. tcgen05 matcher, diagonal form (default): a third of the MMAs of 3b
// =====================================================================================================
// The 3x3 patch correlation is a sum along a diagonal of the pixel Gram matrix G[p, q] = <f_in(p), f_ref(q)>:
//     sim[p, q] = sum_{i, j} G[p + i*w_in + j, q + i*w_ref + j].
// The tensor core only computes the tap-COLUMN sums H[p, q] = sum_i G[p + i*w_in, q + i*w_ref] (K = 3*C: row-shifted
// TMA boxes as in 3b, j = 0 only); the three tap-ROW terms are the same H one row and one column further,
//     sim[p, q] = H[p, q] + H[p+1, q+1] + H[p+2, q+2],
// and the epilogue adds them: in TMEM a lane is a row and a register a column, so H[p+1, q+1] is the neighbouring
// lane's next register -- two warp shuffles and two adds per element.  A warp can only reach its own 32 TMEM lanes, so
// the four 32-lane quarters of the accumulator hold 32-row windows that start 30 rows apart (four 32-row TMA boxes per
// A tile) and only the first 30 lanes of a warp own an output row: 120 rows per tile.  Likewise an N tile covers 254
// output columns with 256 accumulator columns.  Accumulation order differs from 3b (fp32 sums over (i, c) in TMEM, then
// fp32 over j), precision is the same.
constexpr int DG_BN = 256, DG_MV = 120, DG_NV = 254, DG_WROWS = 30;

// Epilogue warps (2..5) of the diagonal-form kernels: per item, sim = H[p, q] + H[p+1, q+1] + H[p+2, q+2] from the
// accumulator (lane = row, register = column: two shuffles per element), scaled by the column's (1/norm, mask),
// running (max, first index) per row, one 64-bit atomicMax per row.
__device__ __forceinline__ void diag_epilogue(const TcParams& prm, const float2* __restrict__ colsb,
                                              unsigned long long* __restrict__ keys, float2* s_sb, uint64_t* tfull,
                                              uint64_t* tempty, uint32_t tmem_base, int warp, int lane) {
    const int q = warp & 3;                     // TMEM lane quarter this warp may access
    const int et = threadIdx.x - 64;            // 0..127
    int it = 0;
    for (int item = blockIdx.x; item < prm.total_items; item += gridDim.x, ++it) {
        const int nt = item % prm.n_tiles;
        const int mt = (item / prm.n_tiles) % prm.m_tiles;
        const int pair = item / (prm.n_tiles * prm.m_tiles);
        const int acc = it & 1;
        const uint32_t acc_phase = (it >> 1) & 1;
        const int n0 = nt * DG_NV;
        const int nout = min(DG_NV, prm.hw_ref - n0);           // output columns of this tile
        int ncols = prm.hw_ref - n0 + 2;
        ncols = ncols > DG_BN ? DG_BN : ((ncols + 31) & ~31);           // accumulator columns the MMAs wrote
        float2* sbuf = s_sb + acc * DG_BN;
        const float2* gsb = colsb + (size_t)pair * prm.colsb_stride + n0;
        for (int i = et; i < ncols; i += 128) sbuf[i] = (i < nout) ? gsb[i] : make_float2(0.f, -INFINITY);
        named_bar_sync(1, 128);
        mbar_wait_backoff(&tfull[acc], acc_phase, 256);
        tc_fence_after();
        float best = -INFINITY;
        int bestn = 0;
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + acc * DG_BN;
        for (int c0 = 0; c0 < ncols; c0 += 32) {
            uint32_t v[32], e[2];
            tmem_ld_32x32(taddr + c0, v);
            if (c0 + 32 < ncols) {
                tmem_ld_32x2(taddr + c0 + 32, e);
            } else {
                e[0] = e[1] = 0u;
            }
            tmem_ld_wait();
#pragma unroll
            for (int k = 0; k < 32; ++k) {
                const uint32_t x1 = (k + 1 < 32) ? v[(k + 1) & 31] : e[0];
                const uint32_t x2 = (k + 2 < 32) ? v[(k + 2) & 31] : e[(k + 2) & 1];
                const float h1 = __uint_as_float(__shfl_down_sync(0xffffffffu, x1, 1));
                const float h2 = __uint_as_float(__shfl_down_sync(0xffffffffu, x2, 2));
                const float sum = (__uint_as_float(v[k]) + h1) + h2;
                const float2 s = sbuf[c0 + k];
                const float val = fmaf(sum, s.x, s.y);
                if (val > best) {
                    best = val;
                    bestn = c0 + k;
                }
            }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tempty[acc]);
        const int m = mt * DG_MV + q * DG_WROWS + lane;
        if (lane < DG_WROWS && m < prm.hw_in && best > -INFINITY)
            atomicMax(keys + (size_t)pair * prm.key_stride + m, pack_key(best, (uint32_t)(n0 + bestn)));
    }
}

template <int NPASS>
struct DgCfg {
    static constexpr int BM = 128, BN = 256, BK = 64;
    static constexpr int MV = 120, NV = 254;             // output rows / columns owned by a tile
    static constexpr int WROWS = 30;                     // row distance of consecutive lane quarters
    static constexpr int A_BYTES = BM * 128, B_BYTES = BN * 128;
    static constexpr int NSPLIT = (NPASS == 1) ? 1 : 2;  // hi only, or hi + lo
    static constexpr int STAGE_BYTES = NSPLIT * (A_BYTES + B_BYTES);
    static constexpr int STAGES = (200 * 1024) / STAGE_BYTES;
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 2 * BN * 8 /*scale,bias*/ + 256 /*barriers*/ + 1024;
};

template <int NPASS>
__global__ void __launch_bounds__(192, 1)
match_diag_kernel(const __grid_constant__ CUtensorMap mapA_hi, const __grid_constant__ CUtensorMap mapA_lo,
                  const __grid_constant__ CUtensorMap mapB_hi, const __grid_constant__ CUtensorMap mapB_lo,
                  const float2* __restrict__ colsb, unsigned long long* __restrict__ keys, const TcParams prm) {
    using Cfg = DgCfg<NPASS>;
    constexpr int BN = Cfg::BN, STAGES = Cfg::STAGES;
    extern __shared__ __align__(1024) uint8_t smem[];
    if ((smem_u32(smem) & 1023u) != 0u) __trap();
    uint8_t* stage_base = smem;
    float2* s_sb = reinterpret_cast<float2*>(smem + STAGES * Cfg::STAGE_BYTES);  // [2][BN]
    uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(s_sb) + 2 * BN * 8);
    uint64_t* full = bars;                 // [STAGES]
    uint64_t* empty = bars + STAGES;       // [STAGES]
    uint64_t* tfull = bars + 2 * STAGES;   // [2]
    uint64_t* tempty = tfull + 2;          // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_slabs = prm.C / Cfg::BK;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&mapA_hi);
        tma_prefetch_desc(&mapB_hi);
        if (NPASS > 1) {
            tma_prefetch_desc(&mapA_lo);
            tma_prefetch_desc(&mapB_lo);
        }
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], 1);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(&tfull[a], 1);
            mbar_init(&tempty[a], 4);  // one arrival per epilogue warp
        }
        fence_mbar_init();
    }
    if (warp == 1) tmem_alloc<512>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ------------------------------------------------------------------ TMA producer
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int item = blockIdx.x; item < prm.total_items; item += gridDim.x) {
                const int nt = item % prm.n_tiles;
                const int mt = (item / prm.n_tiles) % prm.m_tiles;
                const int pair = item / (prm.n_tiles * prm.m_tiles);
                const int img = (pair / prm.in_div) % prm.n_in;
                const int m0 = mt * Cfg::MV, n0 = nt * Cfg::NV;
                for (int slab = 0; slab < n_slabs; ++slab) {
                    for (int ti = 0; ti < 3; ++ti) {
                        const int ra = m0 + ti * prm.w_in, rb = n0 + ti * prm.w_ref;
                        mbar_wait_backoff(&empty[stage], phase ^ 1, 64);
                        uint8_t* sa = stage_base + stage * Cfg::STAGE_BYTES;
                        uint8_t* sbm = sa + Cfg::NSPLIT * Cfg::A_BYTES;
                        mbar_expect_tx(&full[stage], Cfg::STAGE_BYTES);
                        tma_load_3d(sbm, &mapB_hi, &full[stage], slab * Cfg::BK, rb, pair);
                        if (NPASS > 1) tma_load_3d(sbm + Cfg::B_BYTES, &mapB_lo, &full[stage], slab * Cfg::BK, rb, pair);
#pragma unroll
                        for (int wq = 0; wq < 4; ++wq) {
                            tma_load_3d(sa + wq * 4096, &mapA_hi, &full[stage], slab * Cfg::BK, ra + wq * Cfg::WROWS, img);
                            if (NPASS > 1)
                                tma_load_3d(sa + Cfg::A_BYTES + wq * 4096, &mapA_lo, &full[stage], slab * Cfg::BK,
                                            ra + wq * Cfg::WROWS, img);
                        }
                        if (++stage == STAGES) {
                            stage = 0;
                            phase ^= 1;
                        }
                    }
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ------------------------------------------------------------------ MMA issuer
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            int it = 0;
            for (int item = blockIdx.x; item < prm.total_items; item += gridDim.x, ++it) {
                const int nt = item % prm.n_tiles;
                const int acc = it & 1;
                const uint32_t acc_phase = (it >> 1) & 1;
                int ncols = prm.hw_ref - nt * Cfg::NV + 2;     // +2: the look-ahead columns of the last outputs
                ncols = ncols > BN ? BN : ((ncols + 31) & ~31);
                const uint32_t idesc = umma_idesc(MREFSR_IDESC_16, Cfg::BM, ncols);
                mbar_wait_backoff(&tempty[acc], acc_phase ^ 1, 64);
                tc_fence_after();
                const uint32_t tacc = tmem_base + acc * BN;
                uint32_t accumulate = 0;
                for (int slab = 0; slab < n_slabs; ++slab) {
                    for (int ti = 0; ti < 3; ++ti) {
                        mbar_wait(&full[stage], phase);
                        tc_fence_after();
                        const uint32_t sa = smem_u32(stage_base + stage * Cfg::STAGE_BYTES);
                        const uint32_t sb = sa + Cfg::NSPLIT * Cfg::A_BYTES;
#pragma unroll
                        for (int pass = 0; pass < NPASS; ++pass) {
                            // pass 0: hi*hi, pass 1: hi*lo, pass 2: lo*hi
                            const uint32_t a0 = sa + ((pass == 2) ? Cfg::A_BYTES : 0);
                            const uint32_t b0 = sb + ((pass == 1) ? Cfg::B_BYTES : 0);
#pragma unroll
                            for (int kk = 0; kk < 4; ++kk) {
                                umma_f16(tacc, umma_desc_sw128(a0 + kk * 32, 0), umma_desc_sw128(b0 + kk * 32, 0), idesc,
                                         accumulate);
                                accumulate = 1;
                            }
                        }
                        umma_commit(&empty[stage]);
                        if (++stage == STAGES) {
                            stage = 0;
                            phase ^= 1;
                        }
                    }
                }
                umma_commit(&tfull[acc]);
            }
        }
        __syncwarp();
    } else {
        diag_epilogue(prm, colsb, keys, s_sb, tfull, tempty, tmem_base, warp, lane);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc<512>(tmem_base);
    }
}

// ---- 3d. diagonal form with the three tap rows of B taken from ONE shared-memory strip ---------------------------
// match_diag_kernel is bound by the L2 -> shared-memory feed (64 B/clk/SM wanted, ~42 available with all SMs pulling):
// per 64-channel slab it loads the 256 reference rows three times, once per tap row i, each time w_ref rows further.
// Here a stage is a 32-channel slab (64-byte rows, SWIZZLE_64B) whose B side is the strip of 256 + 2*w_ref rows that
// covers all three tap rows; the MMAs of tap row i start i*w_ref rows into it (descriptor start address, swizzle on
// absolute address bits as for the row-shifted descriptors of 3b).  A side: the 4 x 32-row windows of each tap row as
// before.  Bytes per MMA drop 1.6x (39 B/clk/SM).  A stage completes in two halves (tap row 0 + the first 256 strip
// rows; the rest) so that the first MMAs start before the whole stage has landed.  Served when two stages fit
// (w_ref <= 104); wider grids use match_diag_kernel.
constexpr int DS_BK = 32, DS_ROWB = 64;                       // channels / bytes per shared-memory row
constexpr int DS_A_BYTES = 3 * 128 * DS_ROWB;                 // three tap rows x 128 lanes
constexpr int DS_MAX_STAGES = 4;

__device__ __forceinline__ uint64_t umma_desc_sw64(uint32_t saddr) {
    uint64_t d = 0;
    d |= static_cast<uint64_t>((saddr & 0x3FFFF) >> 4);
    d |= static_cast<uint64_t>(1) << 16;
    d |= static_cast<uint64_t>(512 >> 4) << 32;   // 8-row groups 512 B apart
    d |= static_cast<uint64_t>(1) << 46;
    d |= static_cast<uint64_t>(4) << 61;          // SWIZZLE_64B
    return d;
}

template <int NPASS>
__global__ void __launch_bounds__(192, 1)
match_diag_strip_kernel(const __grid_constant__ CUtensorMap mapA_hi, const __grid_constant__ CUtensorMap mapA_lo,
                        const __grid_constant__ CUtensorMap mapB_hi, const __grid_constant__ CUtensorMap mapB_lo,
                        const __grid_constant__ CUtensorMap mapB2_hi, const __grid_constant__ CUtensorMap mapB2_lo,
                        const float2* __restrict__ colsb, unsigned long long* __restrict__ keys, const TcParams prm) {
    constexpr int NSPLIT = (NPASS == 1) ? 1 : 2;
    extern __shared__ __align__(1024) uint8_t smem[];
    if ((smem_u32(smem) & 1023u) != 0u) __trap();
    const int STAGES = prm.stages;
    uint8_t* stage_base = smem;
    float2* s_sb = reinterpret_cast<float2*>(smem + STAGES * prm.stage_bytes);  // [2][BN]
    uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(s_sb) + 2 * DG_BN * 8);
    uint64_t* full0 = bars;                              // [DS_MAX_STAGES] tap row 0 + strip rows 0..255
    uint64_t* full1 = bars + DS_MAX_STAGES;              // [DS_MAX_STAGES] tap rows 1, 2 + the rest of the strip
    uint64_t* empty = bars + 2 * DS_MAX_STAGES;          // [DS_MAX_STAGES]
    uint64_t* tfull = bars + 3 * DS_MAX_STAGES;          // [2]
    uint64_t* tempty = tfull + 2;                        // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_slabs = prm.C / DS_BK;
    const int tail_rows = prm.strip_rows - 256;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&mapA_hi);
        tma_prefetch_desc(&mapB_hi);
        tma_prefetch_desc(&mapB2_hi);
        if (NPASS > 1) {
            tma_prefetch_desc(&mapA_lo);
            tma_prefetch_desc(&mapB_lo);
            tma_prefetch_desc(&mapB2_lo);
        }
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&full0[s], 1);
            mbar_init(&full1[s], 1);
            mbar_init(&empty[s], 1);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(&tfull[a], 1);
            mbar_init(&tempty[a], 4);  // one arrival per epilogue warp
        }
        fence_mbar_init();
    }
    if (warp == 1) tmem_alloc<512>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ------------------------------------------------------------------ TMA producer
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int item = blockIdx.x; item < prm.total_items; item += gridDim.x) {
                const int nt = item % prm.n_tiles;
                const int mt = (item / prm.n_tiles) % prm.m_tiles;
                const int pair = item / (prm.n_tiles * prm.m_tiles);
                const int img = (pair / prm.in_div) % prm.n_in;
                const int m0 = mt * DG_MV, n0 = nt * DG_NV;
                for (int slab = 0; slab < n_slabs; ++slab) {
                    const int c0 = slab * DS_BK;
                    mbar_wait_backoff(&empty[stage], phase ^ 1, 64);
                    uint8_t* sa = stage_base + stage * prm.stage_bytes;
                    uint8_t* sbm = sa + NSPLIT * DS_A_BYTES;
                    // first half: strip rows 0..255 and tap row 0 of A
                    mbar_expect_tx(&full0[stage], NSPLIT * (256 * DS_ROWB + 128 * DS_ROWB));
                    tma_load_3d(sbm, &mapB_hi, &full0[stage], c0, n0, pair);
                    if (NPASS > 1) tma_load_3d(sbm + prm.b_bytes, &mapB_lo, &full0[stage], c0, n0, pair);
#pragma unroll
                    for (int wq = 0; wq < 4; ++wq) {
                        tma_load_3d(sa + wq * 2048, &mapA_hi, &full0[stage], c0, m0 + wq * DG_WROWS, img);
                        if (NPASS > 1)
                            tma_load_3d(sa + DS_A_BYTES + wq * 2048, &mapA_lo, &full0[stage], c0, m0 + wq * DG_WROWS, img);
                    }
                    // second half: the rest of the strip and tap rows 1, 2 of A
                    mbar_expect_tx(&full1[stage], NSPLIT * (tail_rows * DS_ROWB + 2 * 128 * DS_ROWB));
                    tma_load_3d(sbm + 256 * DS_ROWB, &mapB2_hi, &full1[stage], c0, n0 + 256, pair);
                    if (NPASS > 1)
                        tma_load_3d(sbm + prm.b_bytes + 256 * DS_ROWB, &mapB2_lo, &full1[stage], c0, n0 + 256, pair);
#pragma unroll
                    for (int ti = 1; ti < 3; ++ti) {
#pragma unroll
                        for (int wq = 0; wq < 4; ++wq) {
                            const int ra = m0 + ti * prm.w_in + wq * DG_WROWS;
                            tma_load_3d(sa + ti * 8192 + wq * 2048, &mapA_hi, &full1[stage], c0, ra, img);
                            if (NPASS > 1)
                                tma_load_3d(sa + DS_A_BYTES + ti * 8192 + wq * 2048, &mapA_lo, &full1[stage], c0, ra, img);
                        }
                    }
                    if (++stage == STAGES) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ------------------------------------------------------------------ MMA issuer
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            int it = 0;
            for (int item = blockIdx.x; item < prm.total_items; item += gridDim.x, ++it) {
                const int nt = item % prm.n_tiles;
                const int acc = it & 1;
                const uint32_t acc_phase = (it >> 1) & 1;
                int ncols = prm.hw_ref - nt * DG_NV + 2;     // +2: the look-ahead columns of the last outputs
                ncols = ncols > DG_BN ? DG_BN : ((ncols + 31) & ~31);
                const uint32_t idesc = umma_idesc(MREFSR_IDESC_16, 128, ncols);
                mbar_wait_backoff(&tempty[acc], acc_phase ^ 1, 64);
                tc_fence_after();
                const uint32_t tacc = tmem_base + acc * DG_BN;
                uint32_t accumulate = 0;
                for (int slab = 0; slab < n_slabs; ++slab) {
                    const uint32_t sa = smem_u32(stage_base + stage * prm.stage_bytes);
                    const uint32_t sb = sa + NSPLIT * DS_A_BYTES;
#pragma unroll
                    for (int ti = 0; ti < 3; ++ti) {
                        if (ti == 0) mbar_wait(&full0[stage], phase);
                        if (ti == 1) mbar_wait(&full1[stage], phase);
                        if (ti < 2) tc_fence_after();
                        const uint32_t brow = sb + (uint32_t)(ti * prm.w_ref) * DS_ROWB;
#pragma unroll
                        for (int pass = 0; pass < NPASS; ++pass) {
                            // pass 0: hi*hi, pass 1: hi*lo, pass 2: lo*hi
                            const uint32_t a0 = sa + ((pass == 2) ? DS_A_BYTES : 0) + ti * 8192;
                            const uint32_t b0 = brow + ((pass == 1) ? prm.b_bytes : 0);
#pragma unroll
                            for (int kk = 0; kk < 2; ++kk) {
                                umma_f16(tacc, umma_desc_sw64(a0 + kk * 32), umma_desc_sw64(b0 + kk * 32), idesc, accumulate);
                                accumulate = 1;
                            }
                        }
                    }
                    umma_commit(&empty[stage]);
                    if (++stage == STAGES) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
                umma_commit(&tfull[acc]);
            }
        }
        __syncwarp();
    } else {
        diag_epilogue(prm, colsb, keys, s_sb, tfull, tempty, tmem_base, warp, lane);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc<512>(tmem_base);
    }
}

// =====================================================================================================
// host side
// =====================================================================================================
struct MatchPlan {
    int mode, strip, base_offset_mode, diag, bstrip;
    int hw_in, hw_ref, ho, wo, ho_ref, wo_ref;
    int key_stride, colsb_stride;
    size_t off_keys, off_sumsq_in, off_sumsq_ref, off_rownorm, off_colsb, off_a0, off_a1, off_b0, off_b1, total;
};

static bool tc_eligible(int C, int ps, int s_in, int s_ref) { return ps == 3 && s_in == 1 && s_ref == 1 && C % 64 == 0; }

static int make_plan(MatchPlan* pl, int n_in, int n_pairs, int C, int h_in, int w_in, int h_ref, int w_ref, int ps,
                     int s_in, int s_ref, int mode_flags) {
    int mode = mode_flags & 0xff;
    if (mode == MREFSR_MATCH_AUTO) mode = tc_eligible(C, ps, s_in, s_ref) ? MREFSR_MATCH_TC_BF16X3 : MREFSR_MATCH_FP32;
    MREFSR_CHECK(mode >= 1 && mode <= 3, ERR_BAD_ARG, "matcher: unknown mode %d", mode);
    if (mode != MREFSR_MATCH_FP32)
        MREFSR_CHECK(tc_eligible(C, ps, s_in, s_ref), ERR_UNSUPPORTED,
                     "matcher: tcgen05 path needs patch_size 3, strides 1, C %% 64 == 0 (got ps=%d, strides %d/%d, C=%d)",
                     ps, s_in, s_ref, C);
    pl->mode = mode;
    pl->strip = (mode_flags & MREFSR_MATCH_FLAG_NO_STRIP) ? 1 : 3;
    pl->base_offset_mode = (mode_flags & MREFSR_MATCH_FLAG_BASE_OFFSET) ? 1 : 0;
    pl->bstrip = (mode_flags & MREFSR_MATCH_FLAG_NO_BSTRIP) ? 0 : 1;
    pl->diag = (mode_flags & (MREFSR_MATCH_FLAG_NO_DIAG | MREFSR_MATCH_FLAG_NO_STRIP | MREFSR_MATCH_FLAG_BASE_OFFSET)) ? 0 : 1;
    pl->hw_in = h_in * w_in;
    pl->hw_ref = h_ref * w_ref;
    pl->ho = (h_in - ps) / s_in + 1;
    pl->wo = (w_in - ps) / s_in + 1;
    pl->ho_ref = (h_ref - ps) / s_ref + 1;
    pl->wo_ref = (w_ref - ps) / s_ref + 1;
    pl->key_stride = (int)align_up(pl->hw_in, 128);
    pl->colsb_stride = (int)align_up(pl->hw_ref, 256) + 256;
    size_t o = 0;
    auto take = [&](size_t bytes) {
        size_t r = o;
        o += align_up(bytes, 1024);
        return r;
    };
    pl->off_keys = take((size_t)n_pairs * pl->key_stride * 8);
    pl->off_sumsq_in = take((size_t)n_in * pl->hw_in * 4);
    pl->off_sumsq_ref = take((size_t)n_pairs * pl->hw_ref * 4);
    pl->off_rownorm = take((size_t)n_in * pl->hw_in * 4);
    pl->off_colsb = take((size_t)n_pairs * pl->colsb_stride * 8);
    const size_t ein = (size_t)n_in * pl->hw_in * C, eref = (size_t)n_pairs * pl->hw_ref * C;
    if (mode == MREFSR_MATCH_FP32) {
        pl->off_a0 = take(ein * 4);
        pl->off_b0 = take(eref * 4);
        pl->off_a1 = pl->off_b1 = 0;
    } else {
        pl->off_a0 = take(ein * 2);
        pl->off_b0 = take(eref * 2);
        pl->off_a1 = (mode == MREFSR_MATCH_TC_BF16X3) ? take(ein * 2) : 0;
        pl->off_b1 = (mode == MREFSR_MATCH_TC_BF16X3) ? take(eref * 2) : 0;
    }
    pl->total = o;
    return 0;
}

template <int STRIP, int NPASS>
static int launch_tc(const MatchPlan& pl, uint8_t* ws, int n_in, int n_pairs, int in_div, int C, int w_in, int w_ref,
                     cudaStream_t st) {
    using Cfg = TcCfg<STRIP, NPASS>;
    static_assert(Cfg::STAGES >= 2, "need at least a double-buffered pipeline");
    CUtensorMap mA_hi, mA_lo, mB_hi, mB_lo, mB2_hi, mB2_lo;
    const CUtensorMapDataType dt = MREFSR_TMAP_16;
    const CUtensorMapSwizzle sw = CU_TENSOR_MAP_SWIZZLE_128B;
    int rc;
    void* a_hi = ws + pl.off_a0;
    void* b_hi = ws + pl.off_b0;
    void* a_lo = (NPASS > 1) ? ws + pl.off_a1 : a_hi;
    void* b_lo = (NPASS > 1) ? ws + pl.off_b1 : b_hi;
    if ((rc = make_tensor_map_3d(&mA_hi, dt, 2, a_hi, C, pl.hw_in, n_in, 64, Cfg::A_ROWS, sw))) return rc;
    if ((rc = make_tensor_map_3d(&mA_lo, dt, 2, a_lo, C, pl.hw_in, n_in, 64, Cfg::A_ROWS, sw))) return rc;
    if ((rc = make_tensor_map_3d(&mB_hi, dt, 2, b_hi, C, pl.hw_ref, n_pairs, 64, Cfg::BN, sw))) return rc;
    if ((rc = make_tensor_map_3d(&mB_lo, dt, 2, b_lo, C, pl.hw_ref, n_pairs, 64, Cfg::BN, sw))) return rc;
    if ((rc = make_tensor_map_3d(&mB2_hi, dt, 2, b_hi, C, pl.hw_ref, n_pairs, 64, 8, sw))) return rc;
    if ((rc = make_tensor_map_3d(&mB2_lo, dt, 2, b_lo, C, pl.hw_ref, n_pairs, 64, 8, sw))) return rc;
    TcParams prm;
    prm.n_pairs = n_pairs;
    prm.in_div = in_div;
    prm.n_in = n_in;
    prm.C = C;
    // only pixel-linear indices up to the last valid 3x3 patch origin, (h-3)*w + (w-3), need computing: the two
    // bottom rows can never be an origin (the right-hand columns inside the range are masked by colsb / finalize)
    prm.hw_in = (pl.hw_in / w_in - 3) * w_in + (w_in - 2);
    prm.w_in = w_in;
    prm.hw_ref = (pl.hw_ref / w_ref - 3) * w_ref + (w_ref - 2);
    prm.w_ref = w_ref;
    prm.m_tiles = cdiv(prm.hw_in, Cfg::BM);
    prm.n_tiles = cdiv(prm.hw_ref, Cfg::BN);
    prm.total_items = n_pairs * prm.m_tiles * prm.n_tiles;
    prm.key_stride = pl.key_stride;
    prm.colsb_stride = pl.colsb_stride;
    prm.base_offset_mode = pl.base_offset_mode;
    auto kern = match_tc_kernel<STRIP, NPASS>;
    MREFSR_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
    int grid = sm_count();
    if (grid > prm.total_items) grid = prm.total_items;
    kern<<<grid, 192, Cfg::SMEM_BYTES, st>>>(mA_hi, mA_lo, mB_hi, mB_lo, mB2_hi, mB2_lo,
                                              reinterpret_cast<const float2*>(ws + pl.off_colsb),
                                              reinterpret_cast<unsigned long long*>(ws + pl.off_keys), prm);
    MREFSR_LAUNCH_CHECK();
    count_launches(1);
    return 0;
}

template <int NPASS>
static int launch_diag(const MatchPlan& pl, uint8_t* ws, int n_in, int n_pairs, int in_div, int C, int w_in, int w_ref,
                       cudaStream_t st) {
    using Cfg = DgCfg<NPASS>;
    static_assert(Cfg::STAGES >= 2, "need at least a double-buffered pipeline");
    CUtensorMap mA_hi, mA_lo, mB_hi, mB_lo;
    const CUtensorMapDataType dt = MREFSR_TMAP_16;
    const CUtensorMapSwizzle sw = CU_TENSOR_MAP_SWIZZLE_128B;
    int rc;
    void* a_hi = ws + pl.off_a0;
    void* b_hi = ws + pl.off_b0;
    void* a_lo = (NPASS > 1) ? ws + pl.off_a1 : a_hi;
    void* b_lo = (NPASS > 1) ? ws + pl.off_b1 : b_hi;
    if ((rc = make_tensor_map_3d(&mA_hi, dt, 2, a_hi, C, pl.hw_in, n_in, 64, 32, sw))) return rc;
    if ((rc = make_tensor_map_3d(&mA_lo, dt, 2, a_lo, C, pl.hw_in, n_in, 64, 32, sw))) return rc;
    if ((rc = make_tensor_map_3d(&mB_hi, dt, 2, b_hi, C, pl.hw_ref, n_pairs, 64, Cfg::BN, sw))) return rc;
    if ((rc = make_tensor_map_3d(&mB_lo, dt, 2, b_lo, C, pl.hw_ref, n_pairs, 64, Cfg::BN, sw))) return rc;
    TcParams prm;
    prm.n_pairs = n_pairs;
    prm.in_div = in_div;
    prm.n_in = n_in;
    prm.C = C;
    prm.hw_in = (pl.hw_in / w_in - 3) * w_in + (w_in - 2);      // pixel-linear indices up to the last valid origin
    prm.w_in = w_in;
    prm.hw_ref = (pl.hw_ref / w_ref - 3) * w_ref + (w_ref - 2);
    prm.w_ref = w_ref;
    prm.m_tiles = cdiv(prm.hw_in, Cfg::MV);
    prm.n_tiles = cdiv(prm.hw_ref, Cfg::NV);
    prm.total_items = n_pairs * prm.m_tiles * prm.n_tiles;
    prm.key_stride = pl.key_stride;
    prm.colsb_stride = pl.colsb_stride;
    prm.base_offset_mode = 0;
    auto kern = match_diag_kernel<NPASS>;
    MREFSR_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
    int grid = sm_count();
    if (grid > prm.total_items) grid = prm.total_items;
    kern<<<grid, 192, Cfg::SMEM_BYTES, st>>>(mA_hi, mA_lo, mB_hi, mB_lo, reinterpret_cast<const float2*>(ws + pl.off_colsb),
                                              reinterpret_cast<unsigned long long*>(ws + pl.off_keys), prm);
    MREFSR_LAUNCH_CHECK();
    count_launches(1);
    return 0;
}

// stage geometry of match_diag_strip_kernel; false when two stages do not fit
static bool diag_strip_geometry(int npass, int w_ref, TcParams* prm, int* smem_bytes) {
    const int nsplit = (npass == 1) ? 1 : 2;
    const int strip_rows = 256 + (int)align_up((size_t)2 * w_ref, 16);
    if (strip_rows - 256 > 256) return false;
    const int b_bytes = strip_rows * DS_ROWB;
    const int stage_bytes = nsplit * (DS_A_BYTES + b_bytes);
    int stages = (208 * 1024) / stage_bytes;
    if (stages < 2) return false;
    if (stages > DS_MAX_STAGES) stages = DS_MAX_STAGES;
    prm->strip_rows = strip_rows;
    prm->b_bytes = b_bytes;
    prm->stage_bytes = stage_bytes;
    prm->stages = stages;
    *smem_bytes = stages * stage_bytes + 2 * DG_BN * 8 + 256 + 1024;
    return true;
}

template <int NPASS>
static int launch_diag_strip(const MatchPlan& pl, uint8_t* ws, int n_in, int n_pairs, int in_div, int C, int w_in,
                             int w_ref, cudaStream_t st) {
    CUtensorMap mA_hi, mA_lo, mB_hi, mB_lo, mB2_hi, mB2_lo;
    const CUtensorMapDataType dt = MREFSR_TMAP_16;
    const CUtensorMapSwizzle sw = CU_TENSOR_MAP_SWIZZLE_64B;
    TcParams prm;
    int smem_bytes = 0;
    MREFSR_CHECK(diag_strip_geometry(NPASS, w_ref, &prm, &smem_bytes), ERR_UNSUPPORTED, "matcher: strip geometry");
    int rc;
    void* a_hi = ws + pl.off_a0;
    void* b_hi = ws + pl.off_b0;
    void* a_lo = (NPASS > 1) ? ws + pl.off_a1 : a_hi;
    void* b_lo = (NPASS > 1) ? ws + pl.off_b1 : b_hi;
    const int tail = prm.strip_rows - 256;
    if ((rc = make_tensor_map_3d(&mA_hi, dt, 2, a_hi, C, pl.hw_in, n_in, DS_BK, 32, sw))) return rc;
    if ((rc = make_tensor_map_3d(&mA_lo, dt, 2, a_lo, C, pl.hw_in, n_in, DS_BK, 32, sw))) return rc;
    if ((rc = make_tensor_map_3d(&mB_hi, dt, 2, b_hi, C, pl.hw_ref, n_pairs, DS_BK, 256, sw))) return rc;
    if ((rc = make_tensor_map_3d(&mB_lo, dt, 2, b_lo, C, pl.hw_ref, n_pairs, DS_BK, 256, sw))) return rc;
    if ((rc = make_tensor_map_3d(&mB2_hi, dt, 2, b_hi, C, pl.hw_ref, n_pairs, DS_BK, tail, sw))) return rc;
    if ((rc = make_tensor_map_3d(&mB2_lo, dt, 2, b_lo, C, pl.hw_ref, n_pairs, DS_BK, tail, sw))) return rc;
    prm.n_pairs = n_pairs;
    prm.in_div = in_div;
    prm.n_in = n_in;
    prm.C = C;
    prm.hw_in = (pl.hw_in / w_in - 3) * w_in + (w_in - 2);      // pixel-linear indices up to the last valid origin
    prm.w_in = w_in;
    prm.hw_ref = (pl.hw_ref / w_ref - 3) * w_ref + (w_ref - 2);
    prm.w_ref = w_ref;
    prm.m_tiles = cdiv(prm.hw_in, DG_MV);
    prm.n_tiles = cdiv(prm.hw_ref, DG_NV);
    prm.total_items = n_pairs * prm.m_tiles * prm.n_tiles;
    prm.key_stride = pl.key_stride;
    prm.colsb_stride = pl.colsb_stride;
    prm.base_offset_mode = 0;
    auto kern = match_diag_strip_kernel<NPASS>;
    MREFSR_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
    int grid = sm_count();
    if (grid > prm.total_items) grid = prm.total_items;
    kern<<<grid, 192, smem_bytes, st>>>(mA_hi, mA_lo, mB_hi, mB_lo, mB2_hi, mB2_lo,
                                         reinterpret_cast<const float2*>(ws + pl.off_colsb),
                                         reinterpret_cast<unsigned long long*>(ws + pl.off_keys), prm);
    MREFSR_LAUNCH_CHECK();
    count_launches(1);
    return 0;
}

static int run_match(const float* feat_in, const float* feat_ref, int n_in, int n_pairs, int in_div, int C, int h_in,
                     int w_in, int h_ref, int w_ref, int ps, int s_in, int s_ref, int is_norm, int norm_input,
                     int normalize_pixels, int mode_flags, long long* max_idx, float* max_val, void* workspace,
                     size_t ws_bytes, cudaStream_t st) {
    MREFSR_CHECK(feat_in && feat_ref && max_idx && max_val, ERR_BAD_ARG, "matcher: null pointer argument");
    MREFSR_CHECK(n_in > 0 && n_pairs > 0 && in_div > 0 && C > 0, ERR_BAD_ARG, "matcher: bad sizes");
    MREFSR_CHECK(ps >= 1 && s_in >= 1 && s_ref >= 1, ERR_BAD_ARG, "matcher: bad patch_size / stride");
    MREFSR_CHECK(h_in >= ps && w_in >= ps && h_ref >= ps && w_ref >= ps, ERR_BAD_ARG,
                 "matcher: feature map smaller than the patch (%dx%d / %dx%d, patch %d)", h_in, w_in, h_ref, w_ref, ps);
    MatchPlan pl;
    int rc = make_plan(&pl, n_in, n_pairs, C, h_in, w_in, h_ref, w_ref, ps, s_in, s_ref, mode_flags);
    if (rc) return rc;
    MREFSR_CHECK(workspace && ws_bytes >= pl.total, ERR_WORKSPACE, "matcher: workspace too small (%zu < %zu)", ws_bytes,
                 pl.total);
    MREFSR_CHECK((reinterpret_cast<uintptr_t>(workspace) & 1023) == 0, ERR_WORKSPACE,
                 "matcher: workspace must be 1024-byte aligned");
    uint8_t* ws = static_cast<uint8_t*>(workspace);
    MREFSR_CUDA(cudaMemsetAsync(ws + pl.off_keys, 0, (size_t)n_pairs * pl.key_stride * 8, st));

    // 1. prep
    timing_begin(MREFSR_K_MATCH_PREP, st);
    const size_t prep_smem = (size_t)C * 33 * sizeof(float);
    MREFSR_CHECK(prep_smem <= 200 * 1024, ERR_UNSUPPORTED, "matcher: C = %d too large for the prep tile", C);
    auto prep = (C == 256) ? match_prep_kernel<4> : (C == 128) ? match_prep_kernel<2> : (C == 64) ? match_prep_kernel<1>
                                                                                                  : match_prep_kernel<0>;
    if (prep_smem > 48 * 1024)
        MREFSR_CUDA(cudaFuncSetAttribute(prep, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)prep_smem));
    const bool f32 = pl.mode == MREFSR_MATCH_FP32, x3 = pl.mode == MREFSR_MATCH_TC_BF16X3;
    prep<<<dim3(cdiv(pl.hw_in, 32), n_in), 256, prep_smem, st>>>(
        feat_in, C, pl.hw_in, normalize_pixels, f32 ? reinterpret_cast<float*>(ws + pl.off_a0) : nullptr,
        f32 ? nullptr : reinterpret_cast<bf16*>(ws + pl.off_a0), x3 ? reinterpret_cast<bf16*>(ws + pl.off_a1) : nullptr,
        reinterpret_cast<float*>(ws + pl.off_sumsq_in));
    MREFSR_LAUNCH_CHECK();
    prep<<<dim3(cdiv(pl.hw_ref, 32), n_pairs), 256, prep_smem, st>>>(
        feat_ref, C, pl.hw_ref, normalize_pixels, f32 ? reinterpret_cast<float*>(ws + pl.off_b0) : nullptr,
        f32 ? nullptr : reinterpret_cast<bf16*>(ws + pl.off_b0), x3 ? reinterpret_cast<bf16*>(ws + pl.off_b1) : nullptr,
        reinterpret_cast<float*>(ws + pl.off_sumsq_ref));
    MREFSR_LAUNCH_CHECK();
    // 2. norms
    match_norms_kernel<<<dim3(cdiv(pl.hw_in, 256), n_in), 256, 0, st>>>(
        reinterpret_cast<const float*>(ws + pl.off_sumsq_in), n_in, h_in, w_in, ps, s_in, pl.hw_in, 0, norm_input,
        nullptr, reinterpret_cast<float*>(ws + pl.off_rownorm));
    MREFSR_LAUNCH_CHECK();
    match_norms_kernel<<<dim3(cdiv(pl.colsb_stride, 256), n_pairs), 256, 0, st>>>(
        reinterpret_cast<const float*>(ws + pl.off_sumsq_ref), n_pairs, h_ref, w_ref, ps, s_ref, pl.colsb_stride, 1,
        is_norm, reinterpret_cast<float2*>(ws + pl.off_colsb), nullptr);
    MREFSR_LAUNCH_CHECK();
    count_launches(4);
    timing_end(MREFSR_K_MATCH_PREP, st);
    // 3. correlation + arg-max
    timing_begin(MREFSR_K_MATCH_MAIN, st);
    if (f32) {
        dim3 grid(cdiv(pl.ho_ref * pl.wo_ref, ST), cdiv(pl.ho * pl.wo, ST), n_pairs);
        match_simt_kernel<<<grid, 256, 0, st>>>(
            reinterpret_cast<const float*>(ws + pl.off_a0), reinterpret_cast<const float*>(ws + pl.off_b0),
            reinterpret_cast<const float2*>(ws + pl.off_colsb), reinterpret_cast<unsigned long long*>(ws + pl.off_keys),
            in_div, n_in, C, h_in, w_in, h_ref, w_ref, ps, s_in, s_ref, pl.ho, pl.wo, pl.ho_ref, pl.wo_ref,
            pl.key_stride, pl.colsb_stride);
        MREFSR_LAUNCH_CHECK();
        count_launches(1);
    } else {
        TcParams probe;
        int probe_smem = 0;
        if (pl.diag && pl.bstrip && diag_strip_geometry(x3 ? 3 : 1, w_ref, &probe, &probe_smem)) {
            rc = x3 ? launch_diag_strip<3>(pl, ws, n_in, n_pairs, in_div, C, w_in, w_ref, st)
                    : launch_diag_strip<1>(pl, ws, n_in, n_pairs, in_div, C, w_in, w_ref, st);
        } else if (pl.diag) {
            rc = x3 ? launch_diag<3>(pl, ws, n_in, n_pairs, in_div, C, w_in, w_ref, st)
                    : launch_diag<1>(pl, ws, n_in, n_pairs, in_div, C, w_in, w_ref, st);
        } else if (pl.strip == 3) {
            rc = x3 ? launch_tc<3, 3>(pl, ws, n_in, n_pairs, in_div, C, w_in, w_ref, st)
                    : launch_tc<3, 1>(pl, ws, n_in, n_pairs, in_div, C, w_in, w_ref, st);
        } else {
            rc = x3 ? launch_tc<1, 3>(pl, ws, n_in, n_pairs, in_div, C, w_in, w_ref, st)
                    : launch_tc<1, 1>(pl, ws, n_in, n_pairs, in_div, C, w_in, w_ref, st);
        }
        if (rc) return rc;
    }
    timing_end(MREFSR_K_MATCH_MAIN, st);
    // 4. finalize
    const int total = n_pairs * pl.ho * pl.wo;
    match_finalize_kernel<<<cdiv(total, 256), 256, 0, st>>>(
        reinterpret_cast<const unsigned long long*>(ws + pl.off_keys), reinterpret_cast<const float*>(ws + pl.off_rownorm),
        n_pairs, in_div, n_in, h_in, w_in, w_ref, pl.ho, pl.wo, pl.wo_ref, s_in, s_ref, pl.key_stride, max_idx, max_val);
    MREFSR_LAUNCH_CHECK();
    count_launches(1);
    return 0;
}

// =====================================================================================================
// pre-offsets: idx -> flow -> nine zero-filled shifts at scales 1, 2, 4
// (corres_generation_arch.py:30-47, :70-105; arch_util.py:386-410)
//   pre_s[n, 3i+j, Y, X, :] = s * flow[Y/s - i, X/s - j] if Y/s >= i and X/s >= j (flow = 0 on the 2-wide border)
// =====================================================================================================
__global__ void pre_offsets_kernel(const long long* __restrict__ max_idx, int n_img, int h, int w, int s,
                                   float2* __restrict__ out) {
    const int H = h * s, W = w * s;
    const size_t total = (size_t)n_img * 9 * H * W;
    const int hp = h - 2, wp = w - 2;
    for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (size_t)gridDim.x * blockDim.x) {
        const int X = t % W;
        const int Y = (t / W) % H;
        const int k = (t / ((size_t)W * H)) % 9;
        const int img = t / ((size_t)W * H * 9);
        const int y = Y / s - k / 3, x = X / s - k % 3;
        float2 v = make_float2(0.f, 0.f);
        if (y >= 0 && x >= 0 && y < hp && x < wp) {
            const long long idx = max_idx[((size_t)img * hp + y) * wp + x];
            v.x = (float)((int)(idx % wp) - x) * (float)s;
            v.y = (float)((int)(idx / wp) - y) * (float)s;
        }
        out[t] = v;
    }
}

}  // namespace mrefsr

using namespace mrefsr;

extern "C" {

size_t mrefsr_match_workspace_bytes(int n_in, int n_pairs, int C, int h_in, int w_in, int h_ref, int w_ref, int mode) {
    MatchPlan pl;
    // patch 3 / stride 1 sizes the scratch for every mode (other patch sizes only change ho/wo, not the scratch)
    int m = mode & 0xff;
    if (m == MREFSR_MATCH_AUTO) m = (C % 64 == 0) ? MREFSR_MATCH_TC_BF16X3 : MREFSR_MATCH_FP32;
    const int ps = 3;
    if (h_in < ps || w_in < ps || h_ref < ps || w_ref < ps) return 0;
    // the fp32 scratch (4 B/elem) is never smaller than the bf16 hi+lo scratch, so size for the larger of the
    // requested mode and the AUTO fallback
    size_t best = 0;
    for (int mm : {m, (int)MREFSR_MATCH_FP32}) {
        if (mm != MREFSR_MATCH_FP32 && C % 64 != 0) continue;
        if (make_plan(&pl, n_in, n_pairs, C, h_in, w_in, h_ref, w_ref, ps, 1, 1, mm | (mode & ~0xff)) == 0 && pl.total > best)
            best = pl.total;
    }
    return best;
}

int mrefsr_match_plan(int C, int h_in, int w_in, int h_ref, int w_ref, int patch_size, int input_stride, int ref_stride,
                      int mode, int* meta) {
    MREFSR_CHECK(meta, ERR_BAD_ARG, "match_plan: null meta");
    MREFSR_CHECK(patch_size >= 1 && input_stride >= 1 && ref_stride >= 1 && h_in >= patch_size && w_in >= patch_size &&
                     h_ref >= patch_size && w_ref >= patch_size && C > 0,
                 ERR_BAD_ARG, "match_plan: bad sizes");
    MatchPlan pl;
    int rc = make_plan(&pl, 1, 1, C, h_in, w_in, h_ref, w_ref, patch_size, input_stride, ref_stride, mode);
    if (rc) return rc;
    for (int i = 0; i < 10; ++i) meta[i] = 0;
    const int x3 = pl.mode == MREFSR_MATCH_TC_BF16X3;
    meta[9] = pl.mode;
    if (pl.mode == MREFSR_MATCH_FP32) {       // CUDA-core kernel: 64 x 64 tiles over the patch grids
        meta[0] = 0;
        meta[1] = cdiv(pl.ho * pl.wo, ST);
        meta[2] = cdiv(pl.ho_ref * pl.wo_ref, ST);
        meta[3] = ST;
        meta[4] = ST;
        return 0;
    }
    const int rows_in = (pl.hw_in / w_in - 3) * w_in + (w_in - 2), rows_ref = (pl.hw_ref / w_ref - 3) * w_ref + (w_ref - 2);
    meta[5] = rows_in;
    meta[6] = rows_ref;
    TcParams probe;
    int smem = 0;
    if (pl.diag && pl.bstrip && diag_strip_geometry(x3 ? 3 : 1, w_ref, &probe, &smem)) {
        meta[0] = 4;
        meta[1] = cdiv(rows_in, DG_MV);
        meta[2] = cdiv(rows_ref, DG_NV);
        meta[3] = DG_MV;
        meta[4] = DG_NV;
        meta[7] = probe.stages;
        meta[8] = smem;
    } else if (pl.diag) {
        meta[0] = 3;
        meta[1] = cdiv(rows_in, DG_MV);
        meta[2] = cdiv(rows_ref, DG_NV);
        meta[3] = DG_MV;
        meta[4] = DG_NV;
        meta[7] = x3 ? DgCfg<3>::STAGES : DgCfg<1>::STAGES;
        meta[8] = x3 ? DgCfg<3>::SMEM_BYTES : DgCfg<1>::SMEM_BYTES;
    } else {
        meta[0] = pl.strip == 3 ? 2 : 1;
        meta[1] = cdiv(rows_in, 128);
        meta[2] = cdiv(rows_ref, 256);
        meta[3] = 128;
        meta[4] = 256;
        if (pl.strip == 3) {
            meta[7] = x3 ? TcCfg<3, 3>::STAGES : TcCfg<3, 1>::STAGES;
            meta[8] = x3 ? TcCfg<3, 3>::SMEM_BYTES : TcCfg<3, 1>::SMEM_BYTES;
        } else {
            meta[7] = x3 ? TcCfg<1, 3>::STAGES : TcCfg<1, 1>::STAGES;
            meta[8] = x3 ? TcCfg<1, 3>::SMEM_BYTES : TcCfg<1, 1>::SMEM_BYTES;
        }
    }
    return 0;
}

int mrefsr_feature_match_batched(const float* feat_in, const float* feat_ref, int n_in, int n_pairs, int in_div, int C,
                                 int h_in, int w_in, int h_ref, int w_ref, int patch_size, int input_stride,
                                 int ref_stride, int is_norm, int norm_input, int normalize_pixels, int mode,
                                 int64_t* max_idx, float* max_val, void* workspace, size_t workspace_bytes,
                                 void* stream) {
    return run_match(feat_in, feat_ref, n_in, n_pairs, in_div, C, h_in, w_in, h_ref, w_ref, patch_size, input_stride,
                     ref_stride, is_norm, norm_input, normalize_pixels, mode, reinterpret_cast<long long*>(max_idx),
                     max_val, workspace, workspace_bytes, static_cast<cudaStream_t>(stream));
}

int mrefsr_pre_offsets(const int64_t* max_idx, int n, int h, int w, float* out_s1, float* out_s2, float* out_s4,
                       void* stream) {
    MREFSR_CHECK(max_idx && n > 0 && h > 2 && w > 2, ERR_BAD_ARG, "pre_offsets: bad arguments");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    float* outs[3] = {out_s1, out_s2, out_s4};
    const int scales[3] = {1, 2, 4};
    for (int i = 0; i < 3; ++i) {
        if (!outs[i]) continue;
        const size_t total = (size_t)n * 9 * h * scales[i] * w * scales[i];
        int blocks = (int)((total + 255) / 256);
        const int cap = sm_count() * 16;
        if (blocks > cap) blocks = cap;
        ScopedTiming tm(MREFSR_K_GLUE, st);
        pre_offsets_kernel<<<blocks, 256, 0, st>>>(reinterpret_cast<const long long*>(max_idx), n, h, w, scales[i],
                                                   reinterpret_cast<float2*>(outs[i]));
        MREFSR_LAUNCH_CHECK();
        count_launches(1);
    }
    return 0;
}

int mrefsr_feature_match_batched_host(const float* feat_in, const float* feat_ref, int n_in, int n_pairs, int in_div,
                                      int C, int h_in, int w_in, int h_ref, int w_ref, int patch_size, int input_stride,
                                      int ref_stride, int is_norm, int norm_input, int normalize_pixels, int mode,
                                      int64_t* max_idx, float* max_val, void* stream) {
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    MREFSR_CHECK(patch_size >= 1 && input_stride >= 1 && ref_stride >= 1 && h_in >= patch_size && w_in >= patch_size &&
                     h_ref >= patch_size && w_ref >= patch_size,
                 ERR_BAD_ARG, "matcher: bad sizes");
    const size_t ws = mrefsr_match_workspace_bytes(n_in, n_pairs, C, h_in, w_in, h_ref, w_ref, mode);
    const size_t b_in = align_up((size_t)n_in * C * h_in * w_in * 4, 1024);
    const size_t b_ref = align_up((size_t)n_pairs * C * h_ref * w_ref * 4, 1024);
    const int ho = (h_in - patch_size) / input_stride + 1, wo = (w_in - patch_size) / input_stride + 1;
    const size_t b_idx = align_up((size_t)n_pairs * ho * wo * 8, 1024), b_val = align_up((size_t)n_pairs * ho * wo * 4, 1024);
    void* base = nullptr;
    int rc = arena_get(ws + b_in + b_ref + b_idx + b_val, &base);
    if (rc) return rc;
    uint8_t* p = static_cast<uint8_t*>(base);
    uint8_t* d_ws = p;
    float* d_in = reinterpret_cast<float*>(p + align_up(ws, 1024));
    float* d_ref = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(d_in) + b_in);
    long long* d_idx = reinterpret_cast<long long*>(reinterpret_cast<uint8_t*>(d_ref) + b_ref);
    float* d_val = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(d_idx) + b_idx);
    MREFSR_CUDA(cudaMemcpyAsync(d_in, feat_in, (size_t)n_in * C * h_in * w_in * 4, cudaMemcpyHostToDevice, st));
    MREFSR_CUDA(cudaMemcpyAsync(d_ref, feat_ref, (size_t)n_pairs * C * h_ref * w_ref * 4, cudaMemcpyHostToDevice, st));
    rc = run_match(d_in, d_ref, n_in, n_pairs, in_div, C, h_in, w_in, h_ref, w_ref, patch_size, input_stride, ref_stride,
                   is_norm, norm_input, normalize_pixels, mode, d_idx, d_val, d_ws, align_up(ws, 1024), st);
    if (rc) return rc;
    MREFSR_CUDA(cudaMemcpyAsync(max_idx, d_idx, (size_t)n_pairs * ho * wo * 8, cudaMemcpyDeviceToHost, st));
    MREFSR_CUDA(cudaMemcpyAsync(max_val, d_val, (size_t)n_pairs * ho * wo * 4, cudaMemcpyDeviceToHost, st));
    MREFSR_CUDA(cudaStreamSynchronize(st));
    return 0;
}

}  // extern "C"
