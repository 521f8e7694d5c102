// mrefsr_b200/csrc/dcn_win.cu -- DCNv2 forward on tcgen05 with the bilinear corners gathered from SHARED MEMORY.
//
// Same operator as dcn_tc.cu (basicsr/ops/dcn/src/deform_conv_cuda.cpp:490-569 + deform_conv_cuda_kernel.cu:571-633,
// optionally with the DynAgg glue of ref_mrapa_restoration_arch.py:55-68 folded in), same arithmetic bit for bit, but
// different mechanisms for the two streams that bound dcn_tc_split_kernel:
//
//  * corners.  dcn_tc_split_kernel gathers every corner through L1 with 256-bit loads: with C/dg = 8 channels per
//    deform group a sampling point is four scattered 32-byte sectors, 6.7x the algorithmic bytes, at about one sector
//    per clock per SM.  In MRefSR the offsets are a large but spatially COHERENT flow (the matcher's arg-max map: a
//    reference is mostly a translated view) plus a small learned residual, so all nine taps of a patch of output
//    positions sample one compact window of the reference features.  This kernel stages that window in shared memory
//    with ONE bulk tensor copy per (patch, 32-channel slab) -- TMA box {32 channels, wx, wy} of the NHWC input,
//    128-byte swizzle, out-of-image pixels zero-filled by the copy engine, which is exactly the reference's "corners
//    outside the plane contribute 0" -- and reads the corners with LDS.128 (128 B/clk/SM).  Sampling points outside
//    the window (incoherent flow, large learned offsets) take the global-memory path of dcn_tc.cu per item, so the
//    result never depends on the window guess.
//  * offsets / masks.  27 dg planes per output position are the larger half of the algorithmic bytes.  Loaded into
//    registers two tables ahead (dcn_tc_split_kernel, and the first version of this file) they are bound by memory
//    latency x loads in flight: measured on B200, the first version ran at the SAME speed on coherent and random flows
//    (2.87 ms at the large scale), i.e. the corner path was not what it waited for.  Here four loader warps stream
//    them with cp.async (4-byte, position-major coalesced) into a shared-memory ring 6-8 K steps deep: no registers
//    held across the latency, 36-45 KB in flight per SM.
//
//   CTA tile   : 128 output positions (one 16x8 / 8x16 patch or two 8x8 patches) x all Co.  M = 128 per MMA, two TMEM
//                accumulators at every Co <= 256 (the 256-row kernel has one at Co = 256).
//   window     : per patch (px + 2 + 2 mlo) x (py + 2 + 2 mlo) pixels around patch + predicted translation, where
//                the prediction is the flow of the patch centre (fused mode: from the arg-max map; operator mode:
//                the rounded offset of the centre tap of group 0) and mlo = 3 (2 when shared memory is short):
//                residuals in [-mlo, mlo) hit.  Double buffered; reloaded per (tile, slab) by a dedicated warp.
//   K loop     : (32-channel slab) x (tap), as in dcn_tc.cu; A tile [128 x 32] fp32 produced on the SM, B tile by TMA.
//   warps      : 16 gather (thread = (row, 8-channel chunk): raw offset / mask / arg-max words from the ring, sample
//                decode in registers -- no sample table --, 8 LDS.128 or 4 LDG.256, blend, round to tf32, two swizzled
//                16-byte stores; warps 0..3 also drain TMEM), 4 raw loaders (thread = tile row), 1 MMA issuer,
//                1 window loader = 22 warps, 88 registers.
#include <stdlib.h>
#include "dcn_tc.cuh"
#include "../../include/mrefsr_b200.h"

namespace mrefsr {

constexpr int W_BM = 128;                         // rows (output positions) per CTA tile
constexpr int W_A_BYTES = W_BM * 128;             // A stage: 128 rows x 32 fp32
#ifndef MREFSR_WIN_NG
#define MREFSR_WIN_NG 3
#endif
constexpr int W_NG_MAX = MREFSR_WIN_NG;           // producer groups = pipeline stages (fewer when shared memory is short);
                                                  // 3 groups + 2 warps = 14 warps: 4 per SM sub-partition -> 128 registers
constexpr int W_MMA_WARP = W_NG_MAX * 4;          // 4 warps per group: thread = tile row
constexpr int W_WIN_WARP = W_MMA_WARP + 1;
constexpr int W_THREADS = (W_WIN_WARP + 1) * 32;
constexpr int W_SMEM_MAX = 227 * 1024;

__device__ __forceinline__ bool mbar_test_wait(uint64_t* bar, uint32_t parity) {   // never suspends the warp
    uint32_t ok;
    asm volatile(
        "{\n\t"
        ".reg .pred P;\n\t"
        "mbarrier.test_wait.parity.shared::cta.b64 P, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, P;\n\t"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}

// Ablation switches for bottleneck hunting (variant builds only: -DMREFSR_DCN_DEBUG, env MREFSR_DCN_DBG = bit mask;
// results are wrong by design): 1 producers skip corner loads / blend / stores, 2 no MMAs (commits only),
// 4 no offset / mask loads, 8 no weight-tile TMA, 16 no window TMA, 32 no sample decode, 64 no epilogue stores.
#ifdef MREFSR_DCN_DEBUG
#define DBG(bit) (prm.dbg & (bit))
// MREFSR_DCN_TRACE=1: CTA 0 records clock64() at the events of its K steps 180..243 (16 slots per K step)
#define TRACE(ev, kb)                                                                             \
    do {                                                                                          \
        if (prm.trace && blockIdx.x == 0 && (kb) >= 180 && (kb) < 244)                            \
            prm.trace[((kb) - 180) * 16 + (ev)] = clock64();                                      \
    } while (0)
#else
#define DBG(bit) (false)
#define TRACE(ev, kb) do { } while (0)
#endif

// try_wait with a suspend-time hint: the warp is parked by the hardware for up to `ns` instead of spinning through
// the issue slots the working warps need
__device__ __forceinline__ bool mbar_try_wait_hint(uint64_t* bar, uint32_t parity, uint32_t ns) {
    uint32_t ok;
    asm volatile(
        "{\n\t"
        ".reg .pred P;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2, %3;\n\t"
        "selp.u32 %0, 1, 0, P;\n\t"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity), "r"(ns)
        : "memory");
    return ok != 0;
}

__device__ __forceinline__ float4 lds128(uint32_t a) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a));
    return v;
}
// 8 consecutive channels (chunk ch of the slab) of window row `pix`: two 16-byte pieces at the 128-byte-swizzled
// positions the bulk tensor copy wrote them to (window buffers are 1024-byte aligned, so the swizzle phase of a row
// is its index modulo 8)
__device__ __forceinline__ F8 lds_corner(uint32_t win, int pix, int ch) {
    const uint32_t a = win + (uint32_t)pix * 128u + (uint32_t)((((2 * ch) ^ (pix & 7))) << 4);
    const float4 lo = lds128(a), hi = lds128(a ^ 16u);
    F8 r;
    r.v[0] = make_float2(lo.x, lo.y);
    r.v[1] = make_float2(lo.z, lo.w);
    r.v[2] = make_float2(hi.x, hi.y);
    r.v[3] = make_float2(hi.z, hi.w);
    return r;
}

// Window of patch p of a tile: sample b and the image coordinates (wy0, wx0) of its first pixel.  Computed identically
// once per tile by the window loader, which publishes the origins in shared memory.  false: the patch lies beyond the batch (no window).
template <bool FUSED>
__device__ __forceinline__ bool win_patch_geom(const DcnTcParams& prm, int tile, int p, const float* __restrict__ offset,
                                               const long long* __restrict__ max_idx, int& b, int& wy0, int& wx0) {
    const DcnShape& s = prm.s;
    const int st = tile * prm.npatch + p;
    b = st / prm.nsub;
    if (b >= s.B) return false;
    const int rem = st - b * prm.nsub, ty = rem / prm.nsx, tx = rem - ty * prm.nsx;
    const int py0 = ty << prm.ty_log, px0 = tx << prm.tx_log;
    const int cy = min(py0 + (1 << prm.ty_log >> 1), s.Ho - 1), cx = min(px0 + (1 << prm.tx_log >> 1), s.Wo - 1);
    int fy = 0, fx = 0;                           // predicted translation of the patch
    if (FUSED) {
        // flow cell of the centre tap (tap (1,1) reads the cell one up / left, corres_generation_arch.py:74-79)
        const int qy = min(max(cy / prm.flow_scale - 1, 0), prm.hp - 1), qx = min(max(cx / prm.flow_scale - 1, 0), prm.wp - 1);
        const int mi = ldg_early_s32(reinterpret_cast<const int*>(max_idx + ((size_t)b * prm.hp * prm.wp + qy * prm.wp + qx)));
        const int my = (int)__umulhi((unsigned)mi, prm.wp_magic), mx = mi - my * prm.wp;
        fy = (my - qy) * prm.flow_scale;
        fx = (mx - qx) * prm.flow_scale;
    } else {
        const int K = prm.taps, P = prm.P;
        const float* o = offset + (size_t)b * 2 * s.DG * K * P + (size_t)(2 * (K >> 1)) * P + cy * s.Wo + cx;
        fy = (int)rintf(fminf(fmaxf(__ldg(o), -30000.f), 30000.f));      // NaN -> 0
        fx = (int)rintf(fminf(fmaxf(__ldg(o + P), -30000.f), 30000.f));
    }
    wy0 = min(max(py0 - s.ph + fy - prm.mlo, -30000), 30000);
    wx0 = min(max(px0 - s.pw + fx - prm.mlo, -30000), 30000);
    return true;
}

// Drain one finished accumulator tile (128 rows): tcgen05.ld 32x32b, bias add, leaky-ReLU, position-major stores.
__device__ __forceinline__ void win_epilogue_tile(const DcnTcParams& prm, const float* __restrict__ bias,
                                                  uint32_t tmem_base, uint64_t* tempty_bar, int tile, int buf, int warp,
                                                  int lane) {
    const int Co = prm.s.Co, P = prm.P;
    int b = 0, oy = 0, ox = 0;
    const bool ok = dcn_row_coords(prm, tile, warp * 32 + lane, b, oy, ox);
    const int p = ok ? oy * prm.s.Wo + ox : 0;
    if (!ok) b = 0;
    const int bd = prm.dst_group ? (b / prm.dst_group) * prm.dst_stride + prm.dst_offset + b % prm.dst_group : b;
    int k_lo = 0, k_hi = prm.n_outs, Pd = P;      // destination buffers of this row, plane size there
    size_t o_off = prm.out_nhwc ? ((size_t)bd * P + p) * Co : (size_t)bd * Co * P + p;
    if (prm.dst_slab_rows) {
        k_lo = oy / prm.dst_slab_rows;
        k_hi = k_lo + 1;
        Pd = prm.dst_slab_rows * prm.s.Wo;
        o_off = (size_t)bd * Co * Pd + (size_t)(oy - k_lo * prm.dst_slab_rows) * prm.s.Wo + ox;
    }
    const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16) + buf * Co;
#pragma unroll 1
    for (int c0 = 0; c0 < Co; c0 += 8) {
        uint32_t v[8];
        tmem_ld_32x8(taddr + c0, v);
        tmem_ld_wait();
        if (ok && !DBG(64)) {
            float f[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                f[e] = __uint_as_float(v[e]) + (bias ? __ldg(bias + c0 + e) : 0.f);
                f[e] = f[e] > 0.f ? f[e] : f[e] * prm.out_slope;
            }
#pragma unroll 1
            for (int k = k_lo; k < k_hi; ++k) {
                float* o = prm.outs[k] + o_off;
                if (prm.out_nhwc) {
                    *reinterpret_cast<float4*>(o + c0) = make_float4(f[0], f[1], f[2], f[3]);
                    *reinterpret_cast<float4*>(o + c0 + 4) = make_float4(f[4], f[5], f[6], f[7]);
                } else {
#pragma unroll
                    for (int e = 0; e < 8; ++e) o[(size_t)(c0 + e) * Pd] = f[e];
                }
            }
        }
    }
    tc_fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive(tempty_bar);
}

template <int GS>
struct WinRaw {                                   // offset / mask / arg-max words of one tile row for one K step
    float dy[GS], dx[GS], mk[GS];
    int mi;
};

// Shared memory: [NG stages x (A 16 KB + B Co*128)] [2 window buffers] [barriers 256 B] [window origins 4 tiles x 2]
template <bool FUSED, int GS>
__global__ void __launch_bounds__(W_THREADS, 1)
dcn_win_kernel(const __grid_constant__ CUtensorMap mapW, const __grid_constant__ CUtensorMap mapX,
               const float* __restrict__ xt, const float* __restrict__ offset, const float* __restrict__ mask,
               const long long* __restrict__ max_idx, const float* __restrict__ bias,
               const __grid_constant__ DcnTcParams prm) {
    extern __shared__ __align__(1024) uint8_t smem[];
    if ((smem_u32(smem) & 1023u) != 0u) __trap();
    const DcnShape& s = prm.s;
    const int NG = prm.stages;                                         // producer groups, one pipeline stage each
    uint8_t* win = smem + (size_t)NG * prm.stage_bytes;                // two window buffers
    uint64_t* bars = reinterpret_cast<uint64_t*>(win + 2 * (size_t)prm.win_bytes);
    uint64_t* full = bars;                    // [NG]  A produced (4 warps of the owning group) + B landed (expect_tx)
    uint64_t* empty = full + 4;               // [NG]  stage consumed by the MMAs
    uint64_t* tfull = empty + 4;              // [2]   accumulator complete
    uint64_t* tempty = tfull + 2;             // [2]   accumulator drained (4 warps of the draining group)
    uint64_t* win_full = tempty + 2;          // [2]   window landed (expect_tx)
    uint64_t* win_empty = win_full + 2;       // [2]   window released: one arrive per producer warp per K step (4 K)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(win_empty + 2);
    int* worg_ring = reinterpret_cast<int*>(reinterpret_cast<uint8_t*>(bars) + 192);   // [4 tiles][2 patches]

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int Co = s.Co, C = s.C, K = prm.taps, P = prm.P;
    const int nkb_tile = prm.n_slabs * K;

    if (threadIdx.x == 0) {
        tma_prefetch_desc(&mapW);
        tma_prefetch_desc(&mapX);
        for (int i = 0; i < NG; ++i) {
            mbar_init(&full[i], 4 + 1);
            mbar_init(&empty[i], 1);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(&tfull[a], 1);
            mbar_init(&tempty[a], 4);
            mbar_init(&win_full[a], 1);
            mbar_init(&win_empty[a], 4 * K);
        }
        fence_mbar_init();
    }
    if (warp == W_MMA_WARP) tmem_alloc<512>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const int my_tiles = (prm.tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
    const int total_kb = my_tiles * nkb_tile;   // K steps of this CTA, flattened over its tiles: (tile, slab, tap)

    if (warp < W_MMA_WARP) {
        // ------------------------------------------------------------------ producer groups
        // Group g (4 warps, thread = tile row) produces the K steps kb = g (mod NG) into stage g: raw offset / mask
        // words from registers (loaded two of its own K steps ahead) -> sample decode of the GS deform groups of the
        // slab -> per 8-channel chunk 8 LDS.128 (window) or 4 LDG.256 (global), blend, round to tf32, two swizzled
        // 16-byte stores -> arrive.  One barrier round per NG K steps and a whole 128-byte A row per thread, so the
        // fixed cost of a round (three waits, fence, arrive: ~1000 clocks measured) is paid once per row, not once per
        // 32-byte item, and the NG groups overlap each other's latency.  Group it % NG drains accumulator `it`.
        const int g = warp >> 2;
        if (g < NG) {
            const int erow = (warp & 3) * 32 + lane;
            const int my_patch = erow >> prm.sub_log;
            const int pixbase = my_patch * prm.win_rows;
            const uint32_t a_row = (uint32_t)erow * 128u;
            const uint32_t a_swz = (uint32_t)(erow & 7);
            const uint32_t win_u32 = smem_u32(win);
            const unsigned kw_magic = 65536u / (unsigned)s.kw + 1u;    // tap / kw == (tap * kw_magic) >> 16 for tap < 256
            const int dx_elems = C, dy_elems = s.W * C;
            // ---- epilogue duty: tiles it = g, g + NG, ...
            int ep_next = g, c_ti = 0;             // c_ti: tiles whose K steps of this group are all produced
            auto poll_epilogue = [&]() {           // non-blocking; warp-uniform
                if (ep_next < c_ti) {
                    const int buf = ep_next & 1;
                    if (mbar_test_wait(&tfull[buf], (ep_next >> 1) & 1)) {
                        tc_fence_after();
                        win_epilogue_tile(prm, bias, tmem_base, &tempty[buf], (int)blockIdx.x + ep_next * (int)gridDim.x,
                                          buf, warp & 3, lane);
                        ep_next += NG;
                    }
                }
            };
            auto wait_poll = [&](uint64_t* bar, uint32_t parity) {
                while (!mbar_try_wait(bar, parity)) poll_epilogue();
            };
            // ---- load cursor: the K step whose raw words are loaded next, and the row's state in that tile
            int l_kb = g, l_tile = blockIdx.x, l_slab = 0, l_tap = g;
            int n_yx = -1, n_bH = 0, n_qyx = 0, n_idx0 = 0;
            const float* row_off = offset;
            const float* row_msk = mask;
            auto load_rows = [&](int tile) {
                int b = 0, oy = 0, ox = 0;
                n_yx = -1;
                if (dcn_row_coords(prm, tile, erow, b, oy, ox)) {
                    const int p = oy * s.Wo + ox;
                    n_yx = (oy << 16) | ox;
                    n_bH = b * s.H;
                    if (FUSED) {
                        row_off = offset + (size_t)b * 3 * s.DG * K * P + p;
                        n_qyx = ((oy / prm.flow_scale) << 16) | (ox / prm.flow_scale);
                        n_idx0 = b * prm.hp * prm.wp;
                    } else {
                        row_off = offset + (size_t)b * 2 * s.DG * K * P + p;
                        if (mask) row_msk = mask + (size_t)b * s.DG * K * P + p;
                    }
                }
            };
            typedef WinRaw<GS> Raw;
            auto load_raw = [&](Raw& rr) {         // rr <- raw words of K step l_kb; cursor += NG
                if (n_yx >= 0 && !DBG(4)) {
                    const int dg0 = (prm.cdg >= TBK) ? (l_slab * TBK) / prm.cdg : l_slab * (TBK / prm.cdg);
#pragma unroll
                    for (int gi = 0; gi < GS; ++gi) {
                        const int dgi = dg0 + gi;
                        if (!FUSED) {
                            const unsigned o = (unsigned)((dgi * 2 * K + 2 * l_tap) * P);
                            rr.dy[gi] = ldg_early(row_off + o);
                            rr.dx[gi] = ldg_early(row_off + o + (unsigned)P);
                            rr.mk[gi] = mask ? ldg_early(row_msk + (unsigned)((dgi * K + l_tap) * P)) : 1.f;
                        } else {
                            const unsigned o = (unsigned)(2 * (dgi * K + l_tap) * P);
                            rr.dy[gi] = ldg_early(row_off + o);
                            rr.dx[gi] = ldg_early(row_off + o + (unsigned)P);
                            rr.mk[gi] = ldg_early(row_off + (unsigned)((2 * s.DG * K + dgi * K + l_tap) * P));
                        }
                    }
                    if (FUSED) {
                        const int l_ti = (int)(((unsigned)l_tap * kw_magic) >> 16), l_tj = l_tap - l_ti * s.kw;
                        const int fy = (n_qyx >> 16) - l_ti, fx = (n_qyx & 0xffff) - l_tj;
                        if (fy >= 0 && fx >= 0 && fy < prm.hp && fx < prm.wp)
                            rr.mi = ldg_early_s32(reinterpret_cast<const int*>(max_idx + (n_idx0 + fy * prm.wp + fx)));
                    }
                }
                l_kb += NG;
                l_tap += NG;
                if (l_tap >= K) {
                    l_tap -= K;
                    if (++l_slab == prm.n_slabs) {
                        l_slab = 0;
                        l_tile += gridDim.x;
                        if (l_kb < total_kb) load_rows(l_tile);
                    }
                }
            };
            Raw ra, rb;
            ra.mi = rb.mi = 0;
#pragma unroll
            for (int gi = 0; gi < GS; ++gi) ra.dy[gi] = ra.dx[gi] = ra.mk[gi] = rb.dy[gi] = rb.dx[gi] = rb.mk[gi] = 0.f;
            if (g < total_kb) {
                load_rows(l_tile);
                load_raw(ra);
            }
            int c_yx = n_yx, c_bH = n_bH, c_qyx = n_qyx;               // the row's state in the tile being produced
            if (l_kb < total_kb) load_raw(rb);
            int c_slab = 0, c_tap = g, c_wq = 0, waited_wq = -1;       // c_wq: window index (tile, slab) of K step kb
            uint32_t phase = 0;
            bool new_tile = false;

            auto produce = [&](Raw& rr, int kb) {
                const bool tr = (warp & 3) == 0 && lane == 0;
                if (tr) TRACE(0, kb);
                poll_epilogue();
                if (new_tile) {                    // first K step of this group in a new tile (the load cursor is in it)
                    new_tile = false;
                    c_yx = n_yx;
                    c_bH = n_bH;
                    c_qyx = n_qyx;
                }
                if (waited_wq != c_wq) {           // first use of this window by this thread (also orders worg_ring)
                    wait_poll(&win_full[c_wq & 1], (uint32_t)(c_wq >> 1) & 1u);
                    waited_wq = c_wq;
                }
                if (tr) TRACE(1, kb);
                const int worg = worg_ring[(c_ti & 3) * 2 + my_patch];
                const uint32_t wcur = win_u32 + (uint32_t)(c_wq & 1) * (uint32_t)prm.win_bytes;
                const float* xs = xt + c_slab * TBK;
                // ---- per-row constants of the sample decode
                const bool row_ok = c_yx >= 0;
                const int d_ti = (int)(((unsigned)c_tap * kw_magic) >> 16), d_tj = c_tap - d_ti * s.kw;
                const float ybase = (float)((c_yx >> 16) * s.sh - s.ph + d_ti * s.dh);
                const float xbase = (float)((c_yx & 0xffff) * s.sw - s.pw + d_tj * s.dw);
                float fly = 0.f, flx = 0.f;
                if (FUSED) {
                    const int fy = (c_qyx >> 16) - d_ti, fx = (c_qyx & 0xffff) - d_tj;
                    if (row_ok && fy >= 0 && fx >= 0 && fy < prm.hp && fx < prm.wp) {
                        const int my = (int)__umulhi((unsigned)rr.mi, prm.wp_magic), mx = rr.mi - my * prm.wp;
                        fly = (float)((my - fy) * prm.flow_scale);
                        flx = (float)((mx - fx) * prm.flow_scale);
                    }
                }
                const int wy0 = worg >> 16, wx0 = (worg << 16) >> 16;
                if (tr) TRACE(2, kb);
                wait_poll(&empty[g], phase ^ 1);
                if (tr) TRACE(3, kb);
                uint8_t* A = smem + (size_t)g * prm.stage_bytes;
                if ((warp & 3) == 0 && lane == 0) {   // weight tile of this K step (TMA, lands on the same full barrier)
                    if (DBG(8)) {
                        mbar_arrive(&full[g]);
                    } else {
                        mbar_expect_tx(&full[g], Co * 128);
                        tma_load_3d(A + W_A_BYTES, &mapW, &full[g], c_tap * C + c_slab * TBK, 0, 0);
                    }
                }
#pragma unroll
                for (int gi = 0; gi < GS; ++gi) {
                    // ---- sample decode of deform group gi of the slab: same arithmetic as dcn_tc_split_kernel's
                    // (bit for bit) + the window test
                    int bf = ~pixbase;
                    float w0 = 0.f, w1 = 0.f, w2 = 0.f, w3 = 0.f;
                    if (!DBG(32)) {
                        const float y = ybase + (FUSED ? rr.dy[gi] + fly : rr.dy[gi]);
                        const float x = xbase + (FUSED ? rr.dx[gi] + flx : rr.dx[gi]);
                        const bool in = row_ok && y > -1.f && x > -1.f && y < (float)s.H && x < (float)s.W;
                        const float mk = FUSED ? __fdividef(1.f, 1.f + __expf(-rr.mk[gi])) : rr.mk[gi];
                        const float fy0 = floorf(y), fx0 = floorf(x);
                        const int y0 = (int)fy0, x0 = (int)fx0;
                        const float ly = y - fy0, lx = x - fx0, hy = 1.f - ly, hx = 1.f - lx;
                        const bool ty0 = y0 >= 0, ty1 = y0 <= s.H - 2, tx0 = x0 >= 0, tx1 = x0 <= s.W - 2;
                        const int yc = min(max(y0, 0), s.H - 1), xc = min(max(x0, 0), s.W - 1);
                        int base = ((c_bH + yc) * s.W + xc) * C;
                        base |= (tx0 && tx1) ? 1 : 0;
                        base |= (ty0 && ty1) ? 2 : 0;
                        const int ry = y0 - wy0, rx = x0 - wx0;
                        const bool hit = in && ry >= 0 && rx >= 0 && ry <= prm.wy - 2 && rx <= prm.wx - 2;
                        const int pix = pixbase + ry * prm.wx + rx;
                        const float hym = hy * mk, lym = ly * mk;
                        bf = hit ? ~pix : (in ? base : ~pixbase);
                        w0 = (in && ty0 && tx0) ? hym * hx : 0.f;
                        w1 = (in && ty0 && tx1) ? hym * lx : 0.f;
                        w2 = (in && ty1 && tx0) ? lym * hx : 0.f;
                        w3 = (in && ty1 && tx1) ? lym * lx : 0.f;
                    }
                    if (DBG(1)) continue;
                    const float2 p0 = make_float2(w0, w0), p1 = make_float2(w1, w1), p2 = make_float2(w2, w2),
                                 p3 = make_float2(w3, w3);
#pragma unroll
                    for (int cc = 0; cc < 4 / GS; ++cc) {  // the 8-channel chunks of this deform group within the slab
                        const int ch = gi * (4 / GS) + cc;
                        F8 v0, v1, v2, v3;
                        if (bf < 0) {                      // all four corners inside the staged window
                            const int pix = ~bf;
                            v0 = lds_corner(wcur, pix, ch);
                            v1 = lds_corner(wcur, pix + 1, ch);
                            v2 = lds_corner(wcur, pix + prm.wx, ch);
                            v3 = lds_corner(wcur, pix + prm.wx + 1, ch);
                        } else {                           // through L1, as dcn_tc_split_kernel
                            const unsigned i0 = (unsigned)(bf & ~3) + ch * 8;
                            const unsigned i1 = i0 + ((bf & 1) ? dx_elems : 0);
                            const unsigned i2 = i0 + ((bf & 2) ? dy_elems : 0);
                            const unsigned i3 = i2 + ((bf & 1) ? dx_elems : 0);
                            v0 = ldg8(xs + i0);
                            v1 = ldg8(xs + i1);
                            v2 = ldg8(xs + i2);
                            v3 = ldg8(xs + i3);
                        }
                        float2 o[4];
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            float2 a = __fmul2_rn(p0, v0.v[e]);
                            a = __ffma2_rn(p1, v1.v[e], a);
                            a = __ffma2_rn(p2, v2.v[e], a);
                            a = __ffma2_rn(p3, v3.v[e], a);
                            o[e] = make_float2(tf32_round_bits(a.x), tf32_round_bits(a.y));
                        }
                        const uint32_t c0 = a_row + (((uint32_t)(2 * ch) ^ a_swz) << 4);
                        *reinterpret_cast<float4*>(A + c0) = make_float4(o[0].x, o[0].y, o[1].x, o[1].y);
                        *reinterpret_cast<float4*>(A + (c0 ^ 16u)) = make_float4(o[2].x, o[2].y, o[3].x, o[3].y);
                    }
                }
                if (tr) TRACE(4, kb);
                // ---- the raw set is consumed: refill it for this group's K step after next
                if (l_kb < total_kb) load_raw(rr);
                if (tr) TRACE(5, kb);
                fence_proxy_async();
                __syncwarp();
                if (lane == 0) {
                    mbar_arrive(&full[g]);
                    mbar_arrive(&win_empty[c_wq & 1]);     // this warp is done with the window for this K step
                }
                if (tr) TRACE(6, kb);
                phase ^= 1;
                c_tap += NG;
                if (c_tap >= K) {
                    c_tap -= K;
                    ++c_wq;
                    if (++c_slab == prm.n_slabs) {
                        c_slab = 0;
                        ++c_ti;
                        new_tile = true;
                    }
                }
            };
            for (int kb = g; kb < total_kb; kb += NG) {
                produce(ra, kb);           // leaves ra reloading for kb + 2 NG
                Raw t = ra;                // one copy of the produce code (instruction cache): rotate the two raw sets
                ra = rb;
                rb = t;
            }
            // tiles this group did not produce a last K step for still count as produced once the loop is over
            c_ti = my_tiles;
            while (ep_next < my_tiles) {           // drain the remaining accumulators of this group
                const int buf = ep_next & 1;
                mbar_wait_backoff(&tfull[buf], (ep_next >> 1) & 1, 64);
                tc_fence_after();
                win_epilogue_tile(prm, bias, tmem_base, &tempty[buf], (int)blockIdx.x + ep_next * (int)gridDim.x, buf,
                                  warp & 3, lane);
                ep_next += NG;
            }
        }
    } else if (warp == W_MMA_WARP) {
        // ------------------------------------------------------------------ MMA issuer
        if (lane == 0) {
            const uint32_t idesc = umma_idesc(2, 128, Co);
            int stage = 0, it = 0;
            uint32_t phase = 0;
            for (int tile = blockIdx.x; tile < prm.tiles; tile += gridDim.x, ++it) {
                const int buf = it & 1;
                mbar_wait(&tempty[buf], ((it >> 1) & 1) ^ 1);
                tc_fence_after();
                const uint32_t tacc = tmem_base + buf * Co;
                uint32_t accumulate = 0;
                for (int kb = 0; kb < nkb_tile; ++kb) {
                    TRACE(8, it * nkb_tile + kb);
                    mbar_wait(&full[stage], phase);
                    TRACE(9, it * nkb_tile + kb);
                    tc_fence_after();
                    const uint32_t sa = smem_u32(smem + (size_t)stage * prm.stage_bytes);
                    const uint32_t sb = sa + W_A_BYTES;
#pragma unroll
                    for (int kk = 0; kk < 4; ++kk) {
                        if (!DBG(2))
                            umma_tf32(tacc, umma_desc_sw128(sa + kk * 32, 0), umma_desc_sw128(sb + kk * 32, 0), idesc, accumulate);
                        accumulate = 1;
                    }
                    umma_commit(&empty[stage]);
                    TRACE(10, it * nkb_tile + kb);
                    if (++stage == NG) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
                umma_commit(&tfull[buf]);
            }
        }
        __syncwarp();
    } else {
        // ------------------------------------------------------------------ window loader
        if (lane == 0) {
            int wq = 0, ti = 0;
            const uint32_t patch_bytes = (uint32_t)prm.wx * prm.wy * 128u;
            for (int tile = blockIdx.x; tile < prm.tiles; tile += gridDim.x, ++ti) {
                int pb0 = 0, wy00 = 0, wx00 = 0, pb1 = 0, wy01 = 0, wx01 = 0;
                const bool pv0 = win_patch_geom<FUSED>(prm, tile, 0, offset, max_idx, pb0, wy00, wx00);
                const bool pv1 = prm.npatch > 1 && win_patch_geom<FUSED>(prm, tile, 1, offset, max_idx, pb1, wy01, wx01);
                const uint32_t tx_bytes = ((pv0 ? 1u : 0u) + (pv1 ? 1u : 0u)) * patch_bytes;
                for (int slab = 0; slab < prm.n_slabs; ++slab, ++wq) {
                    const int buf = wq & 1;
                    uint8_t* dst = win + (size_t)buf * prm.win_bytes;
                    mbar_wait_backoff(&win_empty[buf], (((uint32_t)wq >> 1) & 1u) ^ 1u, 64);
                    if (slab == 0) {       // published before the tile's first window: ordered by win_full
                        worg_ring[(ti & 3) * 2] = (int)(((unsigned)wy00 << 16) | ((unsigned)wx00 & 0xffffu));
                        worg_ring[(ti & 3) * 2 + 1] = (int)(((unsigned)wy01 << 16) | ((unsigned)wx01 & 0xffffu));
                    }
                    if (DBG(16)) {
                        mbar_arrive(&win_full[buf]);
                        continue;
                    }
                    mbar_expect_tx(&win_full[buf], tx_bytes);
                    if (pv0) tma_load_4d(dst, &mapX, &win_full[buf], slab * TBK, wx00, wy00, pb0);
                    if (pv1) tma_load_4d(dst + (size_t)prm.win_rows * 128, &mapX, &win_full[buf], slab * TBK, wx01, wy01, pb1);
                }
            }
        }
        __syncwarp();
    }
    tc_fence_before();
    __syncthreads();
    if (warp == W_MMA_WARP) {
        tc_fence_after();
        tmem_dealloc<512>(tmem_base);
    }
}

// ---------------------------------------------------------------------------------------------------------
// MREFSR_DCN_WIN=0|1 (default 0) or mrefsr_dcn_window_enable(): route eligible shapes to this kernel.  Opt-in: measured
// on B200 (profiles/r02_dcn_window.md) it is bit-identical to dcn_tc_split_kernel but not faster -- the corner fetch
// stops being the limiter and instruction issue / per-K-step hand-off latency take over.
static int g_win_mode = -1;
static int dcn_win_mode() {
    if (g_win_mode < 0) {
        const char* e = getenv("MREFSR_DCN_WIN");
        g_win_mode = (e && e[0] == '1') ? 1 : 0;
    }
    return g_win_mode;
}
int dcn_win_set_mode(int on) {
    const int prev = dcn_win_mode();
    g_win_mode = on ? 1 : 0;
    return prev;
}

static size_t dcn_win_smem_bytes(const DcnTcParams& prm) {
    return (size_t)prm.stages * prm.stage_bytes + 2 * (size_t)prm.win_bytes + 256;
}

// 128-row tiles as patches; fills the tile / window geometry of prm.  false: shape not served by this kernel.
static bool dcn_win_plan(DcnTcParams& prm) {
    const DcnShape& s = prm.s;
    if (!(s.kh == 3 && s.kw == 3 && s.sh == 1 && s.sw == 1 && s.dh == 1 && s.dw == 1)) return false;
    if (s.H >= 30000 || s.W >= 30000) return false;
    static const int cand[3][2] = {{4, 3}, {3, 4}, {3, 3}};   // log2 (x, y): 16x8, 8x16, two 8x8
    long long best = -1;
    for (int k = 0; k < 3; ++k) {
        const int tx = cand[k][0], ty = cand[k][1];
        const long long nsx = cdiv(s.Wo, 1 << tx), nsy = cdiv(s.Ho, 1 << ty);
        const long long rows = ((long long)s.B * nsx * nsy) << (tx + ty);
        const long long tiles = (rows + W_BM - 1) / W_BM;
        if (best < 0 || tiles < best) {
            best = tiles;
            prm.tx_log = tx;
            prm.ty_log = ty;
            prm.nsx = (int)nsx;
            prm.nsub = (int)(nsx * nsy);
        }
    }
    if (best * W_BM > (long long)prm.total_rows * 140 / 100) return false;   // > 40 % padding rows: not worth it
    prm.tile2d = 1;
    prm.sub_log = prm.tx_log + prm.ty_log;
    prm.subs_log = 7 - prm.sub_log;
    prm.npatch = 1 << prm.subs_log;
    prm.tiles = (int)best;
    prm.stage_bytes = W_A_BYTES + s.Co * 128;
    static const int cfg[5][2] = {{3, 4}, {3, 3}, {2, 3}, {3, 2}, {2, 2}};   // (margin, groups = stages), most wanted first
    for (int k = 0; k < 5; ++k) {
        prm.mlo = cfg[k][0];
        prm.stages = cfg[k][1];
        if (prm.stages > W_NG_MAX) continue;
        prm.rdepth = 0;
        prm.wx = (1 << prm.tx_log) + 2 + 2 * prm.mlo;
        prm.wy = (1 << prm.ty_log) + 2 + 2 * prm.mlo;
        prm.win_rows = (prm.wx * prm.wy + 7) / 8 * 8;
        prm.win_bytes = prm.npatch * prm.win_rows * 128;
        if (dcn_win_smem_bytes(prm) <= (size_t)W_SMEM_MAX) return true;
    }
    return false;
}

// Test hook (host only, no device work): does the window kernel serve a [B, C, H, W] -> [B, Co, H, W] 3x3 / stride 1 /
// pad 1 call with DG deform groups, and with which geometry.  meta[10] = {served, tiles, patch_w, patch_h, patches per
// tile, window_w, window_h, margin, stages, dynamic shared memory bytes}; coords as dcn_tc_tile_plan (128 rows per tile).
int dcn_win_plan_query(int B, int C, int H, int W, int Co, int DG, int* meta, int* coords, size_t max_rows) {
    DcnTcParams prm;
    memset(&prm, 0, sizeof(prm));
    int rc = dcn_make_shape(&prm.s, B, C, H, W, Co, 3, 3, 1, 1, 1, 1, 1, 1, 1, DG);
    if (rc) return rc;
    prm.P = prm.s.Ho * prm.s.Wo;
    prm.total_rows = B * prm.P;
    prm.cdg = C / DG;
    prm.gs = prm.cdg >= TBK ? 1 : TBK / prm.cdg;
    const bool ok = dcn_tc_eligible(prm.s) && dcn_win_plan(prm);
    for (int i = 0; i < 10; ++i) meta[i] = 0;
    if (!ok) return 0;
    meta[0] = 1;
    meta[1] = prm.tiles;
    meta[2] = 1 << prm.tx_log;
    meta[3] = 1 << prm.ty_log;
    meta[4] = prm.npatch;
    meta[5] = prm.wx;
    meta[6] = prm.wy;
    meta[7] = prm.mlo;
    meta[8] = prm.stages;
    meta[9] = (int)dcn_win_smem_bytes(prm);
    if (coords) {
        MREFSR_CHECK((size_t)prm.tiles * W_BM <= max_rows, ERR_BAD_ARG, "window plan: %d rows, room for %zu",
                     prm.tiles * W_BM, max_rows);
        for (int t = 0; t < prm.tiles; ++t)
            for (int r = 0; r < W_BM; ++r) {
                int b = -1, oy = -1, ox = -1;
                if (!dcn_row_coords(prm, t, r, b, oy, ox)) b = oy = ox = -1;
                int* c = coords + ((size_t)t * W_BM + r) * 3;
                c[0] = b;
                c[1] = oy;
                c[2] = ox;
            }
    }
    return 0;
}

// Called by dcn_forward_tc_impl (dcn_tc.cu) with everything but the tile plan filled in.  Returns 1 when the shape is
// not served here (the caller launches dcn_tc_split_kernel), 0 on success, < 0 on error.
int dcn_win_launch(const CUtensorMap& mapW, const float* xt, const float* off, const float* mask,
                   const long long* max_idx, const float* bias, DcnTcParams prm, cudaStream_t st) {
    if (dcn_win_mode() == 0 || !dcn_win_plan(prm)) return 1;
    const DcnShape& s = prm.s;
    prm.nbuf = 2;
#ifdef MREFSR_DCN_DEBUG
    prm.dbg = getenv("MREFSR_DCN_DBG") ? atoi(getenv("MREFSR_DCN_DBG")) : 0;
    prm.trace = nullptr;
    static long long* trace_buf = nullptr;
    if (getenv("MREFSR_DCN_TRACE")) {
        if (!trace_buf) cudaMalloc(&trace_buf, 64 * 16 * sizeof(long long));
        cudaMemsetAsync(trace_buf, 0, 64 * 16 * sizeof(long long), st);
        prm.trace = trace_buf;
    }
#endif
    CUtensorMap mapX;
    int rc = make_tensor_map_4d(&mapX, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, xt, (uint64_t)s.C, (uint64_t)s.W, (uint64_t)s.H,
                                (uint64_t)s.B, TBK, (uint32_t)prm.wx, (uint32_t)prm.wy, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
    const size_t smem = dcn_win_smem_bytes(prm);
    int grid = sm_count();
    if (grid > prm.tiles) grid = prm.tiles;
    ScopedTiming tm(MREFSR_K_DCN_FWD, st);
#define MREFSR_LAUNCH_WIN(F, G)                                                                                     \
    do {                                                                                                            \
        auto kern = dcn_win_kernel<F, G>;                                                                           \
        MREFSR_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));            \
        kern<<<grid, W_THREADS, smem, st>>>(mapW, mapX, xt, off, F ? nullptr : mask, F ? max_idx : nullptr, bias, prm); \
    } while (0)
    if (prm.fused) {
        if (prm.gs == 1) MREFSR_LAUNCH_WIN(true, 1);
        else if (prm.gs == 2) MREFSR_LAUNCH_WIN(true, 2);
        else MREFSR_LAUNCH_WIN(true, 4);
    } else {
        if (prm.gs == 1) MREFSR_LAUNCH_WIN(false, 1);
        else if (prm.gs == 2) MREFSR_LAUNCH_WIN(false, 2);
        else MREFSR_LAUNCH_WIN(false, 4);
    }
#undef MREFSR_LAUNCH_WIN
    MREFSR_LAUNCH_CHECK();
#ifdef MREFSR_DCN_DEBUG
    if (prm.trace) {
        static long long host[64 * 16];
        cudaStreamSynchronize(st);
        cudaMemcpy(host, prm.trace, sizeof(host), cudaMemcpyDeviceToHost);
        long long t0 = host[8] ? host[8] : host[0];
        fprintf(stderr, "TRACE C=%d Co=%d ng=%d  (clocks relative to the MMA thread's step 180; P0 begin, P1 window ok, P2 "
                        "decode consts, P3 stage free, P4 gathered, P5 raw reloaded, P6 arrived | M8 wait full, M9 full ok, M10 committed)\n",
                s.C, s.Co, prm.stages);
        for (int k = 0; k < 64; ++k) {
            fprintf(stderr, "kb %3d g%d:", 180 + k, (180 + k) % prm.stages);
            for (int e = 0; e < 11; ++e)
                if (e < 7 || e >= 8) fprintf(stderr, " %s%d=%6lld", e < 8 ? "P" : "M", e, host[k * 16 + e] ? host[k * 16 + e] - t0 : -1);
            fprintf(stderr, "\n");
        }
    }
#endif
    count_launches(1);
    return 0;
}

}  // namespace mrefsr
