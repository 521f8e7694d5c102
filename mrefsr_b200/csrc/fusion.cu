// mrefsr_b200/csrc/fusion.cu -- multi-reference attention core of MRAPAFusion.
//
// Replaces basicsr/archs/ref_mrapa_restoration_arch.py:321-335: three permute().contiguous() copies, a batched
// [1 x C] . [C x t] matmul per pixel, softmax over t, and a [1 x t] . [t x 2C] matmul per pixel, then a
// permute back.  Here it is one pass over the NCHW tensors exactly as the convolutions left them:
//   CTA = 32 consecutive pixels of one image x 8 warps; lane = pixel (so every global access is a full
//   128-byte line), warp = channel slice.
//   phase 1: per-warp partial logits  l[t] += q[c] * k[t][c]  over its slice of C  -> shared memory
//   phase 2: every thread sums the 8 partials of its pixel, softmax over t in registers
//   phase 3: per-warp slice of the 2C value channels:  out[cv] = sum_t p[t] * v[t][cv]
// HBM-bound: 4*(C + t*C + t*Cv + Cv) bytes per pixel, each byte touched once.
#include "common.cuh"
#include "../../include/mrefsr_b200.h"

namespace mrefsr {

constexpr int FW = 8;  // warps per CTA

template <int TMAX>
__global__ void __launch_bounds__(FW * 32)
mrapa_fwd_kernel(const float* __restrict__ emb_t, const float* __restrict__ emb, const float* __restrict__ ass,
                 float* __restrict__ out, float* __restrict__ prob, int t, int C, int Cv, int HW) {
    __shared__ float part[FW][TMAX][32];
    const int n = blockIdx.y, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int p = blockIdx.x * 32 + lane;
    const bool ok = p < HW;
    const int pc = ok ? p : HW - 1;  // clamp: out-of-range lanes read a valid address and never write
    float l[TMAX];
#pragma unroll
    for (int i = 0; i < TMAX; ++i) l[i] = 0.f;
    const float* q = emb_t + (size_t)n * C * HW + pc;
    const float* k = emb + (size_t)n * t * C * HW + pc;
    const int cper = (C + FW - 1) / FW;
    const int cbeg = warp * cper, cend = min(C, cbeg + cper);
#pragma unroll 4
    for (int c = cbeg; c < cend; ++c) {
        const float qv = __ldg(q + (size_t)c * HW);
#pragma unroll
        for (int i = 0; i < TMAX; ++i)
            if (i < t) l[i] = fmaf(qv, __ldg(k + ((size_t)i * C + c) * HW), l[i]);
    }
#pragma unroll
    for (int i = 0; i < TMAX; ++i) part[warp][i][lane] = l[i];
    __syncthreads();
    float mx = -INFINITY;
#pragma unroll
    for (int i = 0; i < TMAX; ++i) {
        float s = 0.f;
#pragma unroll
        for (int w = 0; w < FW; ++w) s += part[w][i][lane];
        l[i] = s;
        if (i < t) mx = fmaxf(mx, s);
    }
    float den = 0.f;
#pragma unroll
    for (int i = 0; i < TMAX; ++i) {
        l[i] = (i < t) ? expf(l[i] - mx) : 0.f;
        den += l[i];
    }
    const float inv = 1.f / den;
#pragma unroll
    for (int i = 0; i < TMAX; ++i) l[i] *= inv;
    if (prob && warp == 0 && ok) {
#pragma unroll
        for (int i = 0; i < TMAX; ++i)
            if (i < t) prob[((size_t)n * t + i) * HW + p] = l[i];
    }
    const float* v = ass + (size_t)n * t * Cv * HW + pc;
    float* o = out + (size_t)n * Cv * HW + p;
    const int vper = (Cv + FW - 1) / FW;
    const int vbeg = warp * vper, vend = min(Cv, vbeg + vper);
#pragma unroll 4
    for (int c = vbeg; c < vend; ++c) {
        float acc = 0.f;
#pragma unroll
        for (int i = 0; i < TMAX; ++i)
            if (i < t) acc = fmaf(l[i], __ldg(v + ((size_t)i * Cv + c) * HW), acc);
        if (ok) o[(size_t)c * HW] = acc;
    }
}

// Vectorised forward: each lane owns 4 consecutive pixels (one 16-byte load per plane and lane, 512 bytes per
// warp-level load), CTA = 128 pixels x 8 warps.  Needs HW % 4 == 0 and 16-byte aligned tensors.
__device__ __forceinline__ float4 ldcs4(const float* p) { return __ldcs(reinterpret_cast<const float4*>(p)); }
// four consecutive pixels of one channel plane, streamed: fp32 (16 bytes) or bf16 (8 bytes) in memory, fp32 in registers
__device__ __forceinline__ float4 ldcs4(const __nv_bfloat16* p) {
    const uint2 r = __ldcs(reinterpret_cast<const uint2*>(p));
    return make_float4(__uint_as_float(r.x << 16), __uint_as_float(r.x & 0xffff0000u), __uint_as_float(r.y << 16),
                       __uint_as_float(r.y & 0xffff0000u));
}
__device__ __forceinline__ void stcs4(float* p, float4 v) { __stcs(reinterpret_cast<float4*>(p), v); }
__device__ __forceinline__ void stcs4(__nv_bfloat16* p, float4 v) {
    const __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y), b = __floats2bfloat162_rn(v.z, v.w);   // round to nearest even
    uint2 r;
    r.x = *reinterpret_cast<const uint32_t*>(&a);
    r.y = *reinterpret_cast<const uint32_t*>(&b);
    __stcs(reinterpret_cast<uint2*>(p), r);
}

// IO = float: the reference's precision.  IO = __nv_bfloat16: bf16 tensors in and out (half the bytes of this
// HBM-bound kernel), logits / softmax / weighted sum still in fp32.
template <int TMAX, typename IO>   // TMAX == t exactly: no dead accumulators, so 3 CTAs (24 warps) fit per SM
__global__ void __launch_bounds__(FW * 32, 3)
mrapa_fwd_vec4_kernel(const IO* __restrict__ emb_t, const IO* __restrict__ emb, const IO* __restrict__ ass,
                      IO* __restrict__ out, float* __restrict__ prob, int t, int C, int Cv, int HW) {
    __shared__ float4 part[FW][TMAX][32];
    const int n = blockIdx.y, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int p = (blockIdx.x * 32 + lane) * 4;
    const bool ok = p < HW;
    const int pc = ok ? p : 0;
    float4 l[TMAX];
#pragma unroll
    for (int i = 0; i < TMAX; ++i) l[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    const IO* q = emb_t + (size_t)n * C * HW + pc;
    const IO* k = emb + (size_t)n * t * C * HW + pc;
    const int cper = (C + FW - 1) / FW;
    const int cbeg = warp * cper, cend = min(C, cbeg + cper);
#pragma unroll 2
    for (int c = cbeg; c < cend; ++c) {
        const float4 qv = ldcs4(q + (size_t)c * HW);
#pragma unroll
        for (int i = 0; i < TMAX; ++i)
            if (i < t) {
                const float4 kv = ldcs4(k + ((size_t)i * C + c) * HW);
                l[i].x = fmaf(qv.x, kv.x, l[i].x);
                l[i].y = fmaf(qv.y, kv.y, l[i].y);
                l[i].z = fmaf(qv.z, kv.z, l[i].z);
                l[i].w = fmaf(qv.w, kv.w, l[i].w);
            }
    }
#pragma unroll
    for (int i = 0; i < TMAX; ++i) part[warp][i][lane] = l[i];
    __syncthreads();
    float4 mx = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
#pragma unroll
    for (int i = 0; i < TMAX; ++i) {
        float4 sacc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int w = 0; w < FW; ++w) {
            const float4 pv = part[w][i][lane];
            sacc.x += pv.x; sacc.y += pv.y; sacc.z += pv.z; sacc.w += pv.w;
        }
        l[i] = sacc;
        if (i < t) {
            mx.x = fmaxf(mx.x, sacc.x); mx.y = fmaxf(mx.y, sacc.y); mx.z = fmaxf(mx.z, sacc.z); mx.w = fmaxf(mx.w, sacc.w);
        }
    }
    float4 den = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int i = 0; i < TMAX; ++i) {
        if (i < t) {
            l[i].x = expf(l[i].x - mx.x); l[i].y = expf(l[i].y - mx.y); l[i].z = expf(l[i].z - mx.z); l[i].w = expf(l[i].w - mx.w);
        } else {
            l[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        den.x += l[i].x; den.y += l[i].y; den.z += l[i].z; den.w += l[i].w;
    }
    const float4 inv = make_float4(1.f / den.x, 1.f / den.y, 1.f / den.z, 1.f / den.w);
#pragma unroll
    for (int i = 0; i < TMAX; ++i) {
        l[i].x *= inv.x; l[i].y *= inv.y; l[i].z *= inv.z; l[i].w *= inv.w;
    }
    if (prob && warp == 0 && ok) {
#pragma unroll
        for (int i = 0; i < TMAX; ++i)
            if (i < t) *reinterpret_cast<float4*>(prob + ((size_t)n * t + i) * HW + p) = l[i];
    }
    const IO* v = ass + (size_t)n * t * Cv * HW + pc;
    IO* o = out + (size_t)n * Cv * HW + p;
    const int vper = (Cv + FW - 1) / FW;
    const int vbeg = warp * vper, vend = min(Cv, vbeg + vper);
#pragma unroll 2
    for (int c = vbeg; c < vend; ++c) {
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int i = 0; i < TMAX; ++i)
            if (i < t) {
                const float4 vv = ldcs4(v + ((size_t)i * Cv + c) * HW);
                acc.x = fmaf(l[i].x, vv.x, acc.x);
                acc.y = fmaf(l[i].y, vv.y, acc.y);
                acc.z = fmaf(l[i].z, vv.z, acc.z);
                acc.w = fmaf(l[i].w, vv.w, acc.w);
            }
        if (ok) stcs4(o + (size_t)c * HW, acc);
    }
}

// backward:
//   g_ass[t][cv] = p[t] * go[cv];   dp[t] = sum_cv go[cv] * v[t][cv];   dl[t] = p[t] * (dp[t] - sum_s p[s] dp[s])
//   g_q[c] = sum_t dl[t] * k[t][c];   g_k[t][c] = dl[t] * q[c]
template <int TMAX>
__global__ void __launch_bounds__(FW * 32)
mrapa_bwd_kernel(const float* __restrict__ emb_t, const float* __restrict__ emb, const float* __restrict__ ass,
                 const float* __restrict__ prob, const float* __restrict__ gout, float* __restrict__ g_emb_t,
                 float* __restrict__ g_emb, float* __restrict__ g_ass, int t, int C, int Cv, int HW) {
    __shared__ float part[FW][TMAX][32];
    const int n = blockIdx.y, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int p = blockIdx.x * 32 + lane;
    const bool ok = p < HW;
    const int pc = ok ? p : HW - 1;
    float pr[TMAX], dp[TMAX];
#pragma unroll
    for (int i = 0; i < TMAX; ++i) {
        pr[i] = (i < t) ? __ldg(prob + ((size_t)n * t + i) * HW + pc) : 0.f;
        dp[i] = 0.f;
    }
    const float* v = ass + (size_t)n * t * Cv * HW + pc;
    const float* go = gout + (size_t)n * Cv * HW + pc;
    float* gv = g_ass + (size_t)n * t * Cv * HW + p;
    const int vper = (Cv + FW - 1) / FW;
    const int vbeg = warp * vper, vend = min(Cv, vbeg + vper);
#pragma unroll 2
    for (int c = vbeg; c < vend; ++c) {
        const float g = __ldg(go + (size_t)c * HW);
#pragma unroll
        for (int i = 0; i < TMAX; ++i)
            if (i < t) {
                dp[i] = fmaf(g, __ldg(v + ((size_t)i * Cv + c) * HW), dp[i]);
                if (ok) gv[((size_t)i * Cv + c) * HW] = pr[i] * g;
            }
    }
#pragma unroll
    for (int i = 0; i < TMAX; ++i) part[warp][i][lane] = dp[i];
    __syncthreads();
    float dot = 0.f;
#pragma unroll
    for (int i = 0; i < TMAX; ++i) {
        float s = 0.f;
#pragma unroll
        for (int w = 0; w < FW; ++w) s += part[w][i][lane];
        dp[i] = s;
        dot = fmaf(pr[i], s, dot);
    }
#pragma unroll
    for (int i = 0; i < TMAX; ++i) dp[i] = pr[i] * (dp[i] - dot);  // dl
    const float* q = emb_t + (size_t)n * C * HW + pc;
    const float* k = emb + (size_t)n * t * C * HW + pc;
    float* gq = g_emb_t + (size_t)n * C * HW + p;
    float* gk = g_emb + (size_t)n * t * C * HW + p;
    const int cper = (C + FW - 1) / FW;
    const int cbeg = warp * cper, cend = min(C, cbeg + cper);
#pragma unroll 2
    for (int c = cbeg; c < cend; ++c) {
        const float qv = __ldg(q + (size_t)c * HW);
        float acc = 0.f;
#pragma unroll
        for (int i = 0; i < TMAX; ++i)
            if (i < t) {
                acc = fmaf(dp[i], __ldg(k + ((size_t)i * C + c) * HW), acc);
                if (ok) gk[((size_t)i * C + c) * HW] = dp[i] * qv;
            }
        if (ok) gq[(size_t)c * HW] = acc;
    }
}

// The same backward with four consecutive pixels per lane (16-byte accesses, 512 bytes per warp access) and the exact
// reference count as a template parameter (no dead accumulators), like the forward; HW % 4 == 0 and 16-byte aligned
// tensors.  Same arithmetic per element as mrapa_bwd_kernel (the partial sums over the warps' channel slices are added in
// the same order).
template <int T>
__global__ void __launch_bounds__(FW * 32)
mrapa_bwd_vec4_kernel(const float* __restrict__ emb_t, const float* __restrict__ emb, const float* __restrict__ ass,
                      const float* __restrict__ prob, const float* __restrict__ gout, float* __restrict__ g_emb_t,
                      float* __restrict__ g_emb, float* __restrict__ g_ass, int C, int Cv, int HW) {
    __shared__ float4 part[FW][T][32];
    const int n = blockIdx.y, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int p = (blockIdx.x * 32 + lane) * 4;
    const bool ok = p < HW;
    const int pc = ok ? p : HW - 4;
    float4 pr[T], dp[T];
#pragma unroll
    for (int i = 0; i < T; ++i) {
        pr[i] = __ldg(reinterpret_cast<const float4*>(prob + ((size_t)n * T + i) * HW + pc));
        dp[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    const float* v = ass + (size_t)n * T * Cv * HW + pc;
    const float* go = gout + (size_t)n * Cv * HW + pc;
    float* gv = g_ass + (size_t)n * T * Cv * HW + p;
    const int vper = (Cv + FW - 1) / FW;
    const int vbeg = warp * vper, vend = min(Cv, vbeg + vper);
#pragma unroll 2
    for (int c = vbeg; c < vend; ++c) {
        const float4 g = ldcs4(go + (size_t)c * HW);
#pragma unroll
        for (int i = 0; i < T; ++i) {
            const float4 vv = ldcs4(v + ((size_t)i * Cv + c) * HW);
            dp[i].x = fmaf(g.x, vv.x, dp[i].x);
            dp[i].y = fmaf(g.y, vv.y, dp[i].y);
            dp[i].z = fmaf(g.z, vv.z, dp[i].z);
            dp[i].w = fmaf(g.w, vv.w, dp[i].w);
            if (ok) stcs4(gv + ((size_t)i * Cv + c) * HW, make_float4(pr[i].x * g.x, pr[i].y * g.y, pr[i].z * g.z, pr[i].w * g.w));
        }
    }
#pragma unroll
    for (int i = 0; i < T; ++i) part[warp][i][lane] = dp[i];
    __syncthreads();
    float4 dot = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int i = 0; i < T; ++i) {
        float4 sm = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int w = 0; w < FW; ++w) {
            const float4 q4 = part[w][i][lane];
            sm.x += q4.x; sm.y += q4.y; sm.z += q4.z; sm.w += q4.w;
        }
        dp[i] = sm;
        dot.x = fmaf(pr[i].x, sm.x, dot.x);
        dot.y = fmaf(pr[i].y, sm.y, dot.y);
        dot.z = fmaf(pr[i].z, sm.z, dot.z);
        dot.w = fmaf(pr[i].w, sm.w, dot.w);
    }
#pragma unroll
    for (int i = 0; i < T; ++i) {   // dl
        dp[i].x = pr[i].x * (dp[i].x - dot.x);
        dp[i].y = pr[i].y * (dp[i].y - dot.y);
        dp[i].z = pr[i].z * (dp[i].z - dot.z);
        dp[i].w = pr[i].w * (dp[i].w - dot.w);
    }
    const float* q = emb_t + (size_t)n * C * HW + pc;
    const float* k = emb + (size_t)n * T * C * HW + pc;
    float* gq = g_emb_t + (size_t)n * C * HW + p;
    float* gk = g_emb + (size_t)n * T * C * HW + p;
    const int cper = (C + FW - 1) / FW;
    const int cbeg = warp * cper, cend = min(C, cbeg + cper);
#pragma unroll 2
    for (int c = cbeg; c < cend; ++c) {
        const float4 qv = ldcs4(q + (size_t)c * HW);
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int i = 0; i < T; ++i) {
            const float4 kv = ldcs4(k + ((size_t)i * C + c) * HW);
            acc.x = fmaf(dp[i].x, kv.x, acc.x);
            acc.y = fmaf(dp[i].y, kv.y, acc.y);
            acc.z = fmaf(dp[i].z, kv.z, acc.z);
            acc.w = fmaf(dp[i].w, kv.w, acc.w);
            if (ok) stcs4(gk + ((size_t)i * C + c) * HW, make_float4(dp[i].x * qv.x, dp[i].y * qv.y, dp[i].z * qv.z, dp[i].w * qv.w));
        }
        if (ok) stcs4(gq + (size_t)c * HW, acc);
    }
}

static int check_args(int n, int t, int C, int Cv, int h, int w) {
    MREFSR_CHECK(n > 0 && t > 0 && C > 0 && Cv > 0 && h > 0 && w > 0, ERR_BAD_ARG, "mrapa attention: bad sizes");
    MREFSR_CHECK(t <= 16, ERR_UNSUPPORTED, "mrapa attention: at most 16 references are supported (got %d)", t);
    return 0;
}

// ---------------------------------------------------------------------------------------------------------
// Channels-last forward with the three producing convolutions' epilogues folded in (inference):
//   q = prelu(q_raw + bias_q) * q_scale,  k = prelu(k_raw + bias_k),  v = v_raw + bias_v     (conv_emb1 / conv_emb2 /
//   conv_ass of MRAPAFusion, ref_mrapa_restoration_arch.py:293-302, 321-323) are applied on the fly, so the raw cuDNN
//   outputs are read exactly once and never rewritten.  One warp per pixel, lane = V consecutive channels of every
//   32*V-channel group: every access is a contiguous 128*V-byte row segment.  sum_t p_t = 1, so bias_v is added once.
template <int V>
struct VecT;
template <>
struct VecT<2> { using type = float2; };
template <>
struct VecT<4> { using type = float4; };

template <int V>
__device__ __forceinline__ void vload(float (&d)[V], const float* p) {
    const typename VecT<V>::type v = __ldcs(reinterpret_cast<const typename VecT<V>::type*>(p));
    const float* f = reinterpret_cast<const float*>(&v);
#pragma unroll
    for (int e = 0; e < V; ++e) d[e] = f[e];
}
template <int V>
__device__ __forceinline__ void vload_param(float (&d)[V], const float* p, int n, int c, float dflt) {
#pragma unroll
    for (int e = 0; e < V; ++e) d[e] = p ? __ldg(p + (n == 1 ? 0 : c + e)) : dflt;
}

template <int V, int J>   // C = 32 * V * J, Cv = 2 * C
__global__ void __launch_bounds__(256)
mrapa_fwd_nhwc_kernel(const float* __restrict__ q_raw, const float* __restrict__ k_raw, const float* __restrict__ v_raw,
                      const float* __restrict__ bias_q, const float* __restrict__ bias_k,
                      const float* __restrict__ bias_v, const float* __restrict__ slope_q, int slope_q_n,
                      const float* __restrict__ slope_k, int slope_k_n, float q_scale, float* __restrict__ out, int n,
                      int t, long long HW) {
    constexpr int C = 32 * V * J, Cv = 2 * C, JV = 2 * J;
    const int lane = threadIdx.x & 31;
    const long long warp0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    float bq[J][V], bk[J][V], sq[J][V], sk[J][V];
#pragma unroll
    for (int j = 0; j < J; ++j) {
        const int c = j * 32 * V + lane * V;
        vload_param<V>(bq[j], bias_q, 0, c, 0.f);
        vload_param<V>(bk[j], bias_k, 0, c, 0.f);
        vload_param<V>(sq[j], slope_q, slope_q_n, c, 1.f);     // no activation = slope 1
        vload_param<V>(sk[j], slope_k, slope_k_n, c, 1.f);
    }
    for (long long px = warp0; px < (long long)n * HW; px += nwarps) {
        const long long img = px / HW, p = px - img * HW;
        float q[J][V];
#pragma unroll
        for (int j = 0; j < J; ++j) {
            vload<V>(q[j], q_raw + px * C + j * 32 * V + lane * V);
#pragma unroll
            for (int e = 0; e < V; ++e) {
                const float a = q[j][e] + bq[j][e];
                q[j][e] = (a > 0.f ? a : a * sq[j][e]) * q_scale;
            }
        }
        float l[8];
        float mx = -INFINITY;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            l[i] = 0.f;
            if (i < t) {
                const float* kp = k_raw + ((img * t + i) * HW + p) * C + lane * V;
                float part = 0.f;
#pragma unroll
                for (int j = 0; j < J; ++j) {
                    float k[V];
                    vload<V>(k, kp + j * 32 * V);
#pragma unroll
                    for (int e = 0; e < V; ++e) {
                        const float a = k[e] + bk[j][e];
                        part = fmaf(q[j][e], a > 0.f ? a : a * sk[j][e], part);
                    }
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
                l[i] = part;
                mx = fmaxf(mx, part);
            }
        }
        float den = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            l[i] = (i < t) ? expf(l[i] - mx) : 0.f;
            den += l[i];
        }
        const float inv = 1.f / den;
        float acc[JV][V];
#pragma unroll
        for (int j = 0; j < JV; ++j)
#pragma unroll
            for (int e = 0; e < V; ++e) acc[j][e] = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (i < t) {
                const float pi = l[i] * inv;
                const float* vp = v_raw + ((img * t + i) * HW + p) * Cv + lane * V;
#pragma unroll
                for (int j = 0; j < JV; ++j) {
                    float v[V];
                    vload<V>(v, vp + j * 32 * V);
#pragma unroll
                    for (int e = 0; e < V; ++e) acc[j][e] = fmaf(pi, v[e], acc[j][e]);
                }
            }
        }
        float* op = out + px * Cv + lane * V;
#pragma unroll
        for (int j = 0; j < JV; ++j) {
            typename VecT<V>::type o;
            float* f = reinterpret_cast<float*>(&o);
#pragma unroll
            for (int e = 0; e < V; ++e) f[e] = acc[j][e] + (bias_v ? __ldg(bias_v + j * 32 * V + lane * V + e) : 0.f);
            *reinterpret_cast<typename VecT<V>::type*>(op + j * 32 * V) = o;
        }
    }
}

}  // namespace mrefsr

using namespace mrefsr;

extern "C" {

int mrefsr_mrapa_attention_forward(const float* emb_t, const float* emb, const float* ass, float* out, float* prob,
                                   int n, int t, int C, int Cv, int h, int w, void* stream) {
    MREFSR_CHECK(emb_t && emb && ass && out, ERR_BAD_ARG, "mrapa attention forward: null pointer argument");
    int rc = check_args(n, t, C, Cv, h, w);
    if (rc) return rc;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int HW = h * w;
    dim3 grid(cdiv(HW, 32), n);
    ScopedTiming tm(MREFSR_K_FUSION_FWD, st);
    auto al16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
    if (t <= 8 && HW % 4 == 0 && al16(emb_t) && al16(emb) && al16(ass) && al16(out) && (!prob || al16(prob))) {
        const dim3 g4(cdiv(HW, 128), n);
#define MREFSR_FWD4(T) mrapa_fwd_vec4_kernel<T, float><<<g4, FW * 32, 0, st>>>(emb_t, emb, ass, out, prob, t, C, Cv, HW)
        switch (t) {
            case 1: MREFSR_FWD4(1); break;
            case 2: MREFSR_FWD4(2); break;
            case 3: MREFSR_FWD4(3); break;
            case 4: MREFSR_FWD4(4); break;
            case 5: MREFSR_FWD4(5); break;
            case 6: MREFSR_FWD4(6); break;
            case 7: MREFSR_FWD4(7); break;
            default: MREFSR_FWD4(8); break;
        }
#undef MREFSR_FWD4
        MREFSR_LAUNCH_CHECK();
        count_launches(1);
        return 0;
    }
    if (t <= 8)
        mrapa_fwd_kernel<8><<<grid, FW * 32, 0, st>>>(emb_t, emb, ass, out, prob, t, C, Cv, HW);
    else
        mrapa_fwd_kernel<16><<<grid, FW * 32, 0, st>>>(emb_t, emb, ass, out, prob, t, C, Cv, HW);
    MREFSR_LAUNCH_CHECK();
    count_launches(1);
    return 0;
}

int mrefsr_mrapa_attention_forward_bf16(const void* emb_t, const void* emb, const void* ass, void* out, int n, int t, int C,
                                        int Cv, int h, int w, void* stream) {
    MREFSR_CHECK(emb_t && emb && ass && out, ERR_BAD_ARG, "mrapa attention forward (bf16): null pointer argument");
    int rc = check_args(n, t, C, Cv, h, w);
    if (rc) return rc;
    const int HW = h * w;
    auto al8 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 7) == 0; };
    MREFSR_CHECK(t <= 8 && HW % 4 == 0 && al8(emb_t) && al8(emb) && al8(ass) && al8(out), ERR_UNSUPPORTED,
                 "mrapa attention forward (bf16): needs t <= 8, h*w %% 4 == 0 and 8-byte aligned tensors");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const __nv_bfloat16* q = static_cast<const __nv_bfloat16*>(emb_t);
    const __nv_bfloat16* k = static_cast<const __nv_bfloat16*>(emb);
    const __nv_bfloat16* v = static_cast<const __nv_bfloat16*>(ass);
    __nv_bfloat16* o = static_cast<__nv_bfloat16*>(out);
    const dim3 g4(cdiv(HW, 128), n);
    ScopedTiming tm(MREFSR_K_FUSION_FWD, st);
#define MREFSR_FWD4B(T) mrapa_fwd_vec4_kernel<T, __nv_bfloat16><<<g4, FW * 32, 0, st>>>(q, k, v, o, nullptr, t, C, Cv, HW)
    switch (t) {
        case 1: MREFSR_FWD4B(1); break;
        case 2: MREFSR_FWD4B(2); break;
        case 3: MREFSR_FWD4B(3); break;
        case 4: MREFSR_FWD4B(4); break;
        case 5: MREFSR_FWD4B(5); break;
        case 6: MREFSR_FWD4B(6); break;
        case 7: MREFSR_FWD4B(7); break;
        default: MREFSR_FWD4B(8); break;
    }
#undef MREFSR_FWD4B
    MREFSR_LAUNCH_CHECK();
    count_launches(1);
    return 0;
}

int mrefsr_mrapa_attention_backward(const float* emb_t, const float* emb, const float* ass, const float* prob,
                                    const float* grad_out, float* grad_emb_t, float* grad_emb, float* grad_ass, int n,
                                    int t, int C, int Cv, int h, int w, void* stream) {
    MREFSR_CHECK(emb_t && emb && ass && prob && grad_out && grad_emb_t && grad_emb && grad_ass, ERR_BAD_ARG,
                 "mrapa attention backward: null pointer argument");
    int rc = check_args(n, t, C, Cv, h, w);
    if (rc) return rc;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int HW = h * w;
    auto al16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
    if (t <= 8 && HW % 4 == 0 && al16(emb_t) && al16(emb) && al16(ass) && al16(prob) && al16(grad_out) && al16(grad_emb_t) &&
        al16(grad_emb) && al16(grad_ass)) {
        dim3 g4(cdiv(HW, 128), n);
#define MREFSR_BWD4(T) mrapa_bwd_vec4_kernel<T><<<g4, FW * 32, 0, st>>>(emb_t, emb, ass, prob, grad_out, grad_emb_t, grad_emb, grad_ass, C, Cv, HW)
        switch (t) {
            case 1: MREFSR_BWD4(1); break;
            case 2: MREFSR_BWD4(2); break;
            case 3: MREFSR_BWD4(3); break;
            case 4: MREFSR_BWD4(4); break;
            case 5: MREFSR_BWD4(5); break;
            case 6: MREFSR_BWD4(6); break;
            case 7: MREFSR_BWD4(7); break;
            default: MREFSR_BWD4(8); break;
        }
#undef MREFSR_BWD4
        MREFSR_LAUNCH_CHECK();
        count_launches(1);
        return 0;
    }
    dim3 grid(cdiv(HW, 32), n);
    if (t <= 8)
        mrapa_bwd_kernel<8><<<grid, FW * 32, 0, st>>>(emb_t, emb, ass, prob, grad_out, grad_emb_t, grad_emb, grad_ass, t,
                                                      C, Cv, HW);
    else
        mrapa_bwd_kernel<16><<<grid, FW * 32, 0, st>>>(emb_t, emb, ass, prob, grad_out, grad_emb_t, grad_emb, grad_ass,
                                                       t, C, Cv, HW);
    MREFSR_LAUNCH_CHECK();
    count_launches(1);
    return 0;
}

int mrefsr_mrapa_attention_nhwc(const float* q_raw, const float* k_raw, const float* v_raw, const float* bias_q,
                                const float* bias_k, const float* bias_v, const float* slope_q, int slope_q_n,
                                const float* slope_k, int slope_k_n, float q_scale, float* out, int n, int t, int C,
                                int Cv, int h, int w, void* stream) {
    MREFSR_CHECK(q_raw && k_raw && v_raw && out, ERR_BAD_ARG, "mrapa attention nhwc: null pointer argument");
    int rc = check_args(n, t, C, Cv, h, w);
    if (rc) return rc;
    MREFSR_CHECK(t <= 8, ERR_UNSUPPORTED, "mrapa attention nhwc: at most 8 references per call (t=%d)", t);
    MREFSR_CHECK(Cv == 2 * C && (C == 64 || C == 128 || C == 256), ERR_UNSUPPORTED,
                 "mrapa attention nhwc: needs C in {64,128,256} and Cv = 2C (C=%d Cv=%d)", C, Cv);
    MREFSR_CHECK((!slope_q || slope_q_n == 1 || slope_q_n == C) && (!slope_k || slope_k_n == 1 || slope_k_n == C),
                 ERR_BAD_ARG, "mrapa attention nhwc: PReLU weights must have 1 or C entries");
    auto al16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
    MREFSR_CHECK(al16(q_raw) && al16(k_raw) && al16(v_raw) && al16(out), ERR_BAD_ARG,
                 "mrapa attention nhwc: tensors must be 16-byte aligned");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const long long HW = (long long)h * w, warps = (long long)n * HW;
    long long blocks = (warps + 7) / 8;
    if (blocks > 148 * 8 * 4) blocks = 148 * 8 * 4;
    ScopedTiming tm(MREFSR_K_FUSION_FWD, st);
#define MREFSR_NHWC(V, J)                                                                                             \
    mrapa_fwd_nhwc_kernel<V, J><<<(int)blocks, 256, 0, st>>>(q_raw, k_raw, v_raw, bias_q, bias_k, bias_v, slope_q,    \
                                                             slope_q_n, slope_k, slope_k_n, q_scale, out, n, t, HW)
    if (C == 64) MREFSR_NHWC(2, 1);
    else if (C == 128) MREFSR_NHWC(4, 1);
    else MREFSR_NHWC(4, 2);
#undef MREFSR_NHWC
    MREFSR_LAUNCH_CHECK();
    count_launches(1);
    return 0;
}

int mrefsr_mrapa_attention_forward_host(const float* emb_t, const float* emb, const float* ass, float* out, int n, int t,
                                        int C, int Cv, int h, int w, void* stream) {
    int rc = check_args(n, t, C, Cv, h, w);
    if (rc) return rc;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const size_t HW = (size_t)h * w;
    const size_t b_q = align_up(n * C * HW * 4, 1024), b_k = align_up((size_t)n * t * C * HW * 4, 1024),
                 b_v = align_up((size_t)n * t * Cv * HW * 4, 1024), b_o = align_up((size_t)n * Cv * HW * 4, 1024);
    void* base = nullptr;
    rc = arena_get(b_q + b_k + b_v + b_o, &base);
    if (rc) return rc;
    uint8_t* p = static_cast<uint8_t*>(base);
    float *dq = reinterpret_cast<float*>(p), *dk = reinterpret_cast<float*>(p + b_q),
          *dv = reinterpret_cast<float*>(p + b_q + b_k), *d_o = reinterpret_cast<float*>(p + b_q + b_k + b_v);
    MREFSR_CUDA(cudaMemcpyAsync(dq, emb_t, (size_t)n * C * HW * 4, cudaMemcpyHostToDevice, st));
    MREFSR_CUDA(cudaMemcpyAsync(dk, emb, (size_t)n * t * C * HW * 4, cudaMemcpyHostToDevice, st));
    MREFSR_CUDA(cudaMemcpyAsync(dv, ass, (size_t)n * t * Cv * HW * 4, cudaMemcpyHostToDevice, st));
    rc = mrefsr_mrapa_attention_forward(dq, dk, dv, d_o, nullptr, n, t, C, Cv, h, w, st);
    if (rc) return rc;
    MREFSR_CUDA(cudaMemcpyAsync(out, d_o, (size_t)n * Cv * HW * 4, cudaMemcpyDeviceToHost, st));
    MREFSR_CUDA(cudaStreamSynchronize(st));
    return 0;
}

}  // extern "C"
