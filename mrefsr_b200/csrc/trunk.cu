// mrefsr_b200/csrc/trunk.cu -- elementwise glue of the plain-convolution trunk around the hot path, one pass each.
//
// torch runs a convolution's bias add as a separate broadcast kernel (not vectorised: 12.8 ms of a 58.8 ms batch-16
// forward, profiles/r01_full_model.md), then the activation and the residual add as further passes.  Here
//     x[b,c,:] = act(x[b,c,:] + bias[c]) * scale + residual[b,c,:]                      (in place, NCHW planes)
// covers ResidualBlockNoBN (basicsr/archs/arch_util.py:88-117), the offset convolutions and tails of
// DynamicAggregationRestoration (ref_mrapa_restoration_arch.py:140-259), the VGG / extractor conv+ReLU pairs and the
// MRAPAFusion embeddings (conv + PReLU, * C^-0.5; :293-302, :321-323), and
//     refs = refs * sigmoid(attn_mul + b_mul) * 2 + (attn_add + b_add)                   (:341-344)
// is the spatial-attention modulation.  Both are HBM-bound streaming kernels (float4 per lane).
#include "common.cuh"
#include "../../include/mrefsr_b200.h"

namespace mrefsr {

__device__ __forceinline__ float act_apply(float v, int act, float slope) {
    if (act == MREFSR_ACT_LEAKY) return v > 0.f ? v : v * slope;
    if (act == MREFSR_ACT_SIGMOID) return 1.f / (1.f + __expf(-v));
    return v;
}

// grid (ceil(HW4 / (256 * 2)), planes-chunk): plane = b * C + c
template <bool VEC>
__global__ void __launch_bounds__(256)
bias_act_kernel(float* __restrict__ x, const float* __restrict__ bias, const float* __restrict__ slope_dev, int slope_n,
                const float* __restrict__ residual, int planes, int C, int HW, int act, float slope, float scale,
                int res_div, int res_pre) {
    for (int plane = blockIdx.y; plane < planes; plane += gridDim.y) {
        const int c = plane % C;
        const float b = bias ? __ldg(bias + c) : 0.f;
        const float sl = slope_dev ? __ldg(slope_dev + (slope_n == 1 ? 0 : c)) : slope;
        float* xp = x + (size_t)plane * HW;
        const float* rp = residual ? residual + ((size_t)(plane / C / res_div) * C + c) * HW : nullptr;
        if (VEC) {
            const int n4 = HW >> 2;
            for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += gridDim.x * blockDim.x) {
                float4 v = reinterpret_cast<float4*>(xp)[i];
                float4 r = make_float4(0.f, 0.f, 0.f, 0.f);
                if (rp) r = __ldg(reinterpret_cast<const float4*>(rp) + i);
                const float4 pre = res_pre ? r : make_float4(0.f, 0.f, 0.f, 0.f);
                const float4 post = res_pre ? make_float4(0.f, 0.f, 0.f, 0.f) : r;
                v.x = act_apply(v.x + b + pre.x, act, sl) * scale + post.x;
                v.y = act_apply(v.y + b + pre.y, act, sl) * scale + post.y;
                v.z = act_apply(v.z + b + pre.z, act, sl) * scale + post.z;
                v.w = act_apply(v.w + b + pre.w, act, sl) * scale + post.w;
                reinterpret_cast<float4*>(xp)[i] = v;
            }
        } else {
            for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < HW; i += gridDim.x * blockDim.x) {
                const float r = rp ? __ldg(rp + i) : 0.f;
                xp[i] = act_apply(xp[i] + b + (res_pre ? r : 0.f), act, sl) * scale + (res_pre ? 0.f : r);
            }
        }
    }
}

// refs = refs * sigmoid(mul + bias_mul[c]) * 2 + (add + bias_add[c]), in place on refs
template <bool VEC>
__global__ void __launch_bounds__(256)
attn_modulate_kernel(float* __restrict__ refs, const float* __restrict__ mul, const float* __restrict__ add,
                     const float* __restrict__ bias_mul, const float* __restrict__ bias_add, int planes, int C, int HW) {
    for (int plane = blockIdx.y; plane < planes; plane += gridDim.y) {
        const int c = plane % C;
        const float bm = bias_mul ? __ldg(bias_mul + c) : 0.f, ba = bias_add ? __ldg(bias_add + c) : 0.f;
        const size_t o = (size_t)plane * HW;
        if (VEC) {
            const int n4 = HW >> 2;
            for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += gridDim.x * blockDim.x) {
                float4 r = reinterpret_cast<float4*>(refs + o)[i];
                const float4 m = __ldg(reinterpret_cast<const float4*>(mul + o) + i);
                const float4 a = __ldg(reinterpret_cast<const float4*>(add + o) + i);
                r.x = r.x * (2.f / (1.f + __expf(-(m.x + bm)))) + (a.x + ba);
                r.y = r.y * (2.f / (1.f + __expf(-(m.y + bm)))) + (a.y + ba);
                r.z = r.z * (2.f / (1.f + __expf(-(m.z + bm)))) + (a.z + ba);
                r.w = r.w * (2.f / (1.f + __expf(-(m.w + bm)))) + (a.w + ba);
                reinterpret_cast<float4*>(refs + o)[i] = r;
            }
        } else {
            for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < HW; i += gridDim.x * blockDim.x)
                refs[o + i] = refs[o + i] * (2.f / (1.f + __expf(-(mul[o + i] + bm)))) + (add[o + i] + ba);
        }
    }
}

// channels-last variants: x[b, p, c], c fastest; C % 4 == 0 on the float4 path
template <bool VEC>
__global__ void __launch_bounds__(256)
bias_act_nhwc_kernel(float* __restrict__ x, const float* __restrict__ bias, const float* __restrict__ slope_dev,
                     int slope_n, const float* __restrict__ residual, long long total, int C, int act, float slope,
                     float scale, long long per_sample, int res_div, int res_pre) {
    // residual of sample b lives at sample b / res_div: element i -> i - (b - b / res_div) * per_sample
    const long long stride = (long long)gridDim.x * blockDim.x;
    if (VEC) {
        const int c4n = C >> 2;
        for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < (total >> 2); i += stride) {
            const int c = (int)(i % c4n) * 4;
            float4 v = reinterpret_cast<float4*>(x)[i];
            const float4 b = bias ? __ldg(reinterpret_cast<const float4*>(bias + c)) : make_float4(0.f, 0.f, 0.f, 0.f);
            float4 sl = make_float4(slope, slope, slope, slope);
            if (slope_dev) {
                if (slope_n == 1) sl.x = sl.y = sl.z = sl.w = __ldg(slope_dev);
                else sl = __ldg(reinterpret_cast<const float4*>(slope_dev + c));
            }
            float4 r = make_float4(0.f, 0.f, 0.f, 0.f);
            if (residual) {
                long long ri = i;
                if (res_div > 1) {
                    const long long bb = (i * 4) / per_sample;
                    ri = i - (bb - bb / res_div) * (per_sample >> 2);
                }
                r = __ldg(reinterpret_cast<const float4*>(residual) + ri);
            }
            const float4 pre = res_pre ? r : make_float4(0.f, 0.f, 0.f, 0.f);
            const float4 post = res_pre ? make_float4(0.f, 0.f, 0.f, 0.f) : r;
            v.x = act_apply(v.x + b.x + pre.x, act, sl.x) * scale + post.x;
            v.y = act_apply(v.y + b.y + pre.y, act, sl.y) * scale + post.y;
            v.z = act_apply(v.z + b.z + pre.z, act, sl.z) * scale + post.z;
            v.w = act_apply(v.w + b.w + pre.w, act, sl.w) * scale + post.w;
            reinterpret_cast<float4*>(x)[i] = v;
        }
    } else {
        for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
            const int c = (int)(i % C);
            const float sl = slope_dev ? __ldg(slope_dev + (slope_n == 1 ? 0 : c)) : slope;
            float r = 0.f;
            if (residual) {
                const long long bb = i / per_sample;
                r = __ldg(residual + (i - (bb - bb / res_div) * per_sample));
            }
            x[i] = act_apply(x[i] + (bias ? __ldg(bias + c) : 0.f) + (res_pre ? r : 0.f), act, sl) * scale + (res_pre ? 0.f : r);
        }
    }
}

template <bool VEC>
__global__ void __launch_bounds__(256)
attn_modulate_nhwc_kernel(float* __restrict__ refs, const float* __restrict__ mul, const float* __restrict__ add,
                          const float* __restrict__ bias_mul, const float* __restrict__ bias_add, long long total, int C) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    if (VEC) {
        const int c4n = C >> 2;
        for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < (total >> 2); i += stride) {
            const int c = (int)(i % c4n) * 4;
            const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
            const float4 bm = bias_mul ? __ldg(reinterpret_cast<const float4*>(bias_mul + c)) : z;
            const float4 ba = bias_add ? __ldg(reinterpret_cast<const float4*>(bias_add + c)) : z;
            float4 r = reinterpret_cast<float4*>(refs)[i];
            const float4 m = __ldg(reinterpret_cast<const float4*>(mul) + i);
            const float4 a = __ldg(reinterpret_cast<const float4*>(add) + i);
            r.x = r.x * (2.f / (1.f + __expf(-(m.x + bm.x)))) + (a.x + ba.x);
            r.y = r.y * (2.f / (1.f + __expf(-(m.y + bm.y)))) + (a.y + ba.y);
            r.z = r.z * (2.f / (1.f + __expf(-(m.z + bm.z)))) + (a.z + ba.z);
            r.w = r.w * (2.f / (1.f + __expf(-(m.w + bm.w)))) + (a.w + ba.w);
            reinterpret_cast<float4*>(refs)[i] = r;
        }
    } else {
        for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
            const int c = (int)(i % C);
            refs[i] = refs[i] * (2.f / (1.f + __expf(-(mul[i] + (bias_mul ? bias_mul[c] : 0.f))))) +
                      (add[i] + (bias_add ? bias_add[c] : 0.f));
        }
    }
}

// Dense layout conversion, 32 positions x 128 channels per CTA through shared memory, both directions at streaming
// rate (torch's strided copy_ runs these at ~1/4 of it).  grid (ceil(HW/32), ceil(C/128), B); C % 4 == 0.
__global__ void __launch_bounds__(256)
nhwc_to_nchw_kernel(const float* __restrict__ src, float* __restrict__ dst, const float* __restrict__ bias, int C, int HW) {
    __shared__ float tile[128][33];
    const int b = blockIdx.z, p0 = blockIdx.x * 32, c0 = blockIdx.y * 128;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const float* s = src + (size_t)b * C * HW;
    float* d = dst + (size_t)b * C * HW;
#pragma unroll
    for (int j = 0; j < 4; ++j) {        // loads: one float4 (4 channels) per lane, 512 contiguous bytes per warp
        const int pl = warp * 4 + j, pp = p0 + pl, c = c0 + lane * 4;
        if (pp < HW && c < C) {
            const float4 v = __ldcs(reinterpret_cast<const float4*>(s + (size_t)pp * C + c));
            tile[lane * 4][pl] = v.x;
            tile[lane * 4 + 1][pl] = v.y;
            tile[lane * 4 + 2][pl] = v.z;
            tile[lane * 4 + 3][pl] = v.w;
        }
    }
    __syncthreads();
    const int p = p0 + lane;
#pragma unroll
    for (int i = 0; i < 16; ++i) {       // stores: lane = position, 128-byte coalesced rows of one channel plane
        const int cl = warp + 8 * i, c = c0 + cl;
        if (c < C && p < HW) d[(size_t)c * HW + p] = tile[cl][lane] + (bias ? __ldg(bias + c) : 0.f);
    }
}

__global__ void __launch_bounds__(256)
nchw_to_nhwc_conv_kernel(const float* __restrict__ src, float* __restrict__ dst, const float* __restrict__ bias, int C, int HW) {
    __shared__ float tile[128][33];
    const int b = blockIdx.z, p0 = blockIdx.x * 32, c0 = blockIdx.y * 128;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const float* s = src + (size_t)b * C * HW;
    float* d = dst + (size_t)b * C * HW;
    const int p = p0 + lane;
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        const int cl = warp + 8 * i, c = c0 + cl;
        tile[cl][lane] = (c < C && p < HW) ? __ldcs(s + (size_t)c * HW + p) + (bias ? __ldg(bias + c) : 0.f) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int pl = warp * 4 + j, pp = p0 + pl, c = c0 + lane * 4;
        if (pp < HW && c < C)
            *reinterpret_cast<float4*>(d + (size_t)pp * C + c) =
                make_float4(tile[lane * 4][pl], tile[lane * 4 + 1][pl], tile[lane * 4 + 2][pl], tile[lane * 4 + 3][pl]);
    }
}

// The same two conversions with a dtype change on the way (training under bf16 autocast: the plain convolutions produce
// and consume bf16 channels-last tensors, the alignment kernels fp32 planes): one pass instead of torch's cast followed
// by the transpose.  bf16 side: 4 channels = 8 bytes per lane.
// gate != NULL (same shape / layout as src): dst = src * (gate > 0 ? 1 : slope) -- the backward of a leaky ReLU whose output
// is `gate`, folded into the conversion of the incoming gradient.  CT = channels per CTA tile (128, or 64 so that 64-channel
// tensors -- the large scale -- keep every lane busy): CT / 4 lanes cover a pixel's channels, 8 bytes each.
template <int CT>
__global__ void __launch_bounds__(256)
nhwc_bf16_to_nchw_f32_kernel(const __nv_bfloat16* __restrict__ src, float* __restrict__ dst,
                             const __nv_bfloat16* __restrict__ gate, float slope, int C, int HW) {
    constexpr int LPP = CT / 4, PPW = 32 / LPP, IT = 32 / (8 * PPW);
    __shared__ float tile[CT][33];
    const int b = blockIdx.z, p0 = blockIdx.x * 32, c0 = blockIdx.y * CT;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const __nv_bfloat16* s = src + (size_t)b * C * HW;
    float* d = dst + (size_t)b * C * HW;
    const int cq = (lane % LPP) * 4;
#pragma unroll
    for (int j = 0; j < IT; ++j) {
        const int pl = (warp * IT + j) * PPW + lane / LPP, pp = p0 + pl, c = c0 + cq;
        if (pp < HW && c < C) {
            const uint2 v = __ldcs(reinterpret_cast<const uint2*>(s + (size_t)pp * C + c));
            float f0 = __uint_as_float(v.x << 16), f1 = __uint_as_float(v.x & 0xffff0000u);
            float f2 = __uint_as_float(v.y << 16), f3 = __uint_as_float(v.y & 0xffff0000u);
            if (gate) {
                const uint2 g = __ldcs(reinterpret_cast<const uint2*>(gate + (size_t)b * C * HW + (size_t)pp * C + c));
                // bf16 sign / zero test on the raw bits: positive and non-zero <=> (bits & 0x7fff) != 0 and sign clear
                f0 = ((g.x & 0x8000u) || !(g.x & 0x7fffu)) ? f0 * slope : f0;
                f1 = ((g.x & 0x80000000u) || !(g.x & 0x7fff0000u)) ? f1 * slope : f1;
                f2 = ((g.y & 0x8000u) || !(g.y & 0x7fffu)) ? f2 * slope : f2;
                f3 = ((g.y & 0x80000000u) || !(g.y & 0x7fff0000u)) ? f3 * slope : f3;
            }
            tile[cq][pl] = f0;
            tile[cq + 1][pl] = f1;
            tile[cq + 2][pl] = f2;
            tile[cq + 3][pl] = f3;
        }
    }
    __syncthreads();
    const int p = p0 + lane;
#pragma unroll
    for (int i = 0; i < CT / 8; ++i) {
        const int cl = warp + 8 * i, c = c0 + cl;
        if (c < C && p < HW) d[(size_t)c * HW + p] = tile[cl][lane];
    }
}

// slope != 1: a leaky ReLU applied on the way (dst = lrelu(src))
template <int CT>
__global__ void __launch_bounds__(256)
nchw_f32_to_nhwc_bf16_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, float slope, int C, int HW) {
    constexpr int LPP = CT / 4, PPW = 32 / LPP, IT = 32 / (8 * PPW);
    __shared__ float tile[CT][33];
    const int b = blockIdx.z, p0 = blockIdx.x * 32, c0 = blockIdx.y * CT;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const float* s = src + (size_t)b * C * HW;
    __nv_bfloat16* d = dst + (size_t)b * C * HW;
    const int p = p0 + lane;
#pragma unroll
    for (int i = 0; i < CT / 8; ++i) {
        const int cl = warp + 8 * i, c = c0 + cl;
        float v = (c < C && p < HW) ? __ldcs(s + (size_t)c * HW + p) : 0.f;
        tile[cl][lane] = v > 0.f ? v : v * slope;
    }
    __syncthreads();
    const int cq = (lane % LPP) * 4;
#pragma unroll
    for (int j = 0; j < IT; ++j) {
        const int pl = (warp * IT + j) * PPW + lane / LPP, pp = p0 + pl, c = c0 + cq;
        if (pp < HW && c < C) {
            const __nv_bfloat162 lo = __floats2bfloat162_rn(tile[cq][pl], tile[cq + 1][pl]);
            const __nv_bfloat162 hi = __floats2bfloat162_rn(tile[cq + 2][pl], tile[cq + 3][pl]);
            uint2 o;
            o.x = *reinterpret_cast<const uint32_t*>(&lo);
            o.y = *reinterpret_cast<const uint32_t*>(&hi);
            *reinterpret_cast<uint2*>(d + (size_t)pp * C + c) = o;
        }
    }
}

// 2x2 / stride 2 max pooling on channels-last data (VGG pool1 / pool2, vgg_arch.py:141-161): float4 over channels,
// every access a contiguous row segment.  torch's max_pool_forward_nhwc takes 1.7 ms for the six calls of a batch-16
// forward; this is the streaming rate.
__global__ void __launch_bounds__(256)
maxpool2_nhwc_kernel(const float* __restrict__ src, float* __restrict__ dst, long long total4, int C4, int Wo, int Ho, int W) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total4; i += (long long)gridDim.x * blockDim.x) {
        const int c4 = (int)(i % C4);
        long long r = i / C4;
        const int ox = (int)(r % Wo);
        r /= Wo;
        const int oy = (int)(r % Ho);
        const long long b = r / Ho;
        const float4* s = reinterpret_cast<const float4*>(src) + ((b * (2 * Ho) + 2 * oy) * (long long)W + 2 * ox) * C4 + c4;
        const float4 a = __ldcs(s), bb = __ldcs(s + C4), c = __ldcs(s + (long long)W * C4), d = __ldcs(s + (long long)W * C4 + C4);
        float4 o;
        o.x = fmaxf(fmaxf(a.x, bb.x), fmaxf(c.x, d.x));
        o.y = fmaxf(fmaxf(a.y, bb.y), fmaxf(c.y, d.y));
        o.z = fmaxf(fmaxf(a.z, bb.z), fmaxf(c.z, d.z));
        o.w = fmaxf(fmaxf(a.w, bb.w), fmaxf(c.w, d.w));
        reinterpret_cast<float4*>(dst)[i] = o;
    }
}

static bool al16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

// ---------------------------------------------------------------------------------------------------------
// Training path of the same epilogue (channels-last activations, fp32 or bf16 -- the config-5 step runs the plain
// convolutions under bf16 autocast).  torch computes a convolution's bias gradient as a separate reduction over
// grad_output (0.09 ms per [12, 64, 160, 160] tensor, 10 ms of a 102 ms step) after a separate activation-backward
// pass; here  grad_in = grad_out * act'(y) * scale  and  grad_bias = sum_rows grad_in  are one pass, and the forward
//     y = act(x + bias[c]) * scale + residual
// is one in-place pass over the convolution's (bias-free) output.  Rows = B*H*W pixels, C channels fastest.
// A thread owns one 16-byte channel vector (4 fp32 / 8 bf16) and walks rows; per-channel sums go through shared
// memory to one partial row per CTA, a second tiny kernel adds the partial rows in a fixed order (deterministic).
template <typename T>
struct TrainVec;
template <>
struct TrainVec<float> {
    static constexpr int W = 4;
    __device__ static void load(const float* p, float (&v)[4]) {
        const float4 t = *reinterpret_cast<const float4*>(p);
        v[0] = t.x, v[1] = t.y, v[2] = t.z, v[3] = t.w;
    }
    __device__ static void store(float* p, const float (&v)[4]) {
        *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
    }
};
template <>
struct TrainVec<__nv_bfloat16> {
    static constexpr int W = 8;
    __device__ static void load(const __nv_bfloat16* p, float (&v)[8]) {
        const uint4 t = *reinterpret_cast<const uint4*>(p);
        const uint32_t w[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            v[2 * i] = __uint_as_float(w[i] << 16);
            v[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
        }
    }
    __device__ static void store(__nv_bfloat16* p, const float (&v)[8]) {
        uint32_t w[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const __nv_bfloat162 h = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
            w[i] = *reinterpret_cast<const uint32_t*>(&h);
        }
        *reinterpret_cast<uint4*>(p) = make_uint4(w[0], w[1], w[2], w[3]);
    }
};

template <typename T>
__global__ void __launch_bounds__(256)
bias_act_train_fwd_kernel(T* __restrict__ x, const float* __restrict__ bias, const T* __restrict__ residual, long long nvec,
                          int cvec, int act, float slope, float scale) {
    constexpr int W = TrainVec<T>::W;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += stride) {
        const int c = (int)(i % cvec) * W;
        float v[W], r[W];
        TrainVec<T>::load(x + i * W, v);
        if (residual) TrainVec<T>::load(residual + i * W, r);
#pragma unroll
        for (int k = 0; k < W; ++k) {
            float t = v[k] + (bias ? __ldg(bias + c + k) : 0.f);
            if (act == MREFSR_ACT_LEAKY) t = t > 0.f ? t : t * slope;
            v[k] = t * scale + (residual ? r[k] : 0.f);
        }
        TrainVec<T>::store(x + i * W, v);
    }
}

// grad_in may be NULL (no activation and scale 1: grad_in == grad_out, only the bias gradient is wanted)
template <typename T>
__global__ void __launch_bounds__(256)
bias_act_train_bwd_kernel(const T* __restrict__ gout, const T* __restrict__ y, T* __restrict__ gin,
                          float* __restrict__ partial, long long rows, int C, int act, float slope, float scale) {
    constexpr int W = TrainVec<T>::W;
    __shared__ float red[256 * W];
    const int cvec = C / W, rpb = 256 / cvec;          // rows per CTA per iteration; threads beyond rpb * cvec idle
    const int cv = threadIdx.x % cvec, r0 = threadIdx.x / cvec;
    float acc[W];
#pragma unroll
    for (int k = 0; k < W; ++k) acc[k] = 0.f;
    for (long long row = (long long)blockIdx.x * rpb + r0; r0 < rpb && row < rows; row += (long long)gridDim.x * rpb) {
        const long long o = row * C + (long long)cv * W;
        float g[W], yy[W];
        TrainVec<T>::load(gout + o, g);
        if (act == MREFSR_ACT_LEAKY) {
            TrainVec<T>::load(y + o, yy);
#pragma unroll
            for (int k = 0; k < W; ++k) g[k] = (yy[k] > 0.f ? g[k] : g[k] * slope);
        }
#pragma unroll
        for (int k = 0; k < W; ++k) {
            g[k] *= scale;
            acc[k] += g[k];
        }
        if (gin) TrainVec<T>::store(gin + o, g);
    }
#pragma unroll
    for (int k = 0; k < W; ++k) red[threadIdx.x * W + k] = acc[k];
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += 256) {
        const int v = c / W, k = c - v * W;
        float sum = 0.f;
        for (int r = 0; r < rpb; ++r) sum += red[(r * cvec + v) * W + k];
        partial[(size_t)blockIdx.x * C + c] = sum;
    }
}

// one warp per channel: lanes stride over the partial rows, then a shuffle tree -- a fixed order (deterministic)
__global__ void bias_grad_sum_kernel(const float* __restrict__ partial, float* __restrict__ gbias, int nblocks, int C) {
    const int c = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (c >= C) return;
    float sum = 0.f;
    for (int b = lane; b < nblocks; b += 32) sum += partial[(size_t)b * C + c];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    if (lane == 0) gbias[c] = sum;
}

static bool train_shape_ok(int C, int dtype) {
    const int W = dtype ? 8 : 4;
    return C % W == 0 && C / W <= 256;
}

}  // namespace mrefsr

using namespace mrefsr;

extern "C" {

int mrefsr_bias_act(float* x, const float* bias, const float* slope_dev, int slope_n, const float* residual,
                    int res_div, int res_pre, int B, int C, int HW, int channels_last, int act, float slope,
                    float scale, void* stream) {
    MREFSR_CHECK(res_div >= 1 && (!residual || B % res_div == 0), ERR_BAD_ARG, "bias_act: batch %d not a multiple of res_div %d", B, res_div);
    MREFSR_CHECK(x, ERR_BAD_ARG, "bias_act: null tensor");
    MREFSR_CHECK(B > 0 && C > 0 && HW > 0, ERR_BAD_ARG, "bias_act: bad shape B=%d C=%d HW=%d", B, C, HW);
    MREFSR_CHECK(act == MREFSR_ACT_NONE || act == MREFSR_ACT_LEAKY || act == MREFSR_ACT_SIGMOID, ERR_BAD_ARG,
                 "bias_act: unknown activation %d", act);
    MREFSR_CHECK(!slope_dev || slope_n == 1 || slope_n == C, ERR_BAD_ARG, "bias_act: slope tensor has %d entries, C=%d",
                 slope_n, C);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const long long planes_ll = (long long)B * C;
    MREFSR_CHECK(planes_ll < (1ll << 31), ERR_BAD_ARG, "bias_act: too many planes");
    const int planes = (int)planes_ll;
    ScopedTiming tm(MREFSR_K_GLUE, st);
    if (channels_last) {
        const long long total = planes_ll * HW;
        const bool v4 = C % 4 == 0 && al16(x) && (!residual || al16(residual)) && (!bias || al16(bias)) &&
                        (!slope_dev || slope_n == 1 || al16(slope_dev));
        const long long work = v4 ? total / 4 : total;
        const int blocks = (int)((work + 256 * 4 - 1) / (256 * 4) < 1 ? 1 : (work + 256 * 4 - 1) / (256 * 4) > 148 * 32 ? 148 * 32 : (work + 256 * 4 - 1) / (256 * 4));
        if (v4)
            bias_act_nhwc_kernel<true><<<blocks, 256, 0, st>>>(x, bias, slope_dev, slope_n, residual, total, C, act, slope, scale, (long long)C * HW, res_div, res_pre);
        else
            bias_act_nhwc_kernel<false><<<blocks, 256, 0, st>>>(x, bias, slope_dev, slope_n, residual, total, C, act, slope, scale, (long long)C * HW, res_div, res_pre);
        MREFSR_LAUNCH_CHECK();
        count_launches(1);
        return 0;
    }
    const bool vec = HW % 4 == 0 && al16(x) && (!residual || al16(residual));
    const int per = vec ? HW / 4 : HW;
    dim3 grid(cdiv(per, 256 * 2) < 1 ? 1 : cdiv(per, 256 * 2), planes < 32768 ? planes : 32768);
    if (vec)
        bias_act_kernel<true><<<grid, 256, 0, st>>>(x, bias, slope_dev, slope_n, residual, planes, C, HW, act, slope, scale, res_div, res_pre);
    else
        bias_act_kernel<false><<<grid, 256, 0, st>>>(x, bias, slope_dev, slope_n, residual, planes, C, HW, act, slope, scale, res_div, res_pre);
    MREFSR_LAUNCH_CHECK();
    count_launches(1);
    return 0;
}

int mrefsr_attn_modulate(float* refs, const float* attn_mul, const float* attn_add, const float* bias_mul,
                         const float* bias_add, int B, int C, int HW, int channels_last, void* stream) {
    MREFSR_CHECK(refs && attn_mul && attn_add, ERR_BAD_ARG, "attn_modulate: null tensor");
    MREFSR_CHECK(B > 0 && C > 0 && HW > 0, ERR_BAD_ARG, "attn_modulate: bad shape B=%d C=%d HW=%d", B, C, HW);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const long long planes_ll = (long long)B * C;
    MREFSR_CHECK(planes_ll < (1ll << 31), ERR_BAD_ARG, "attn_modulate: too many planes");
    const int planes = (int)planes_ll;
    ScopedTiming tm(MREFSR_K_GLUE, st);
    if (channels_last) {
        const long long total = planes_ll * HW;
        const bool v4 = C % 4 == 0 && al16(refs) && al16(attn_mul) && al16(attn_add) && (!bias_mul || al16(bias_mul)) &&
                        (!bias_add || al16(bias_add));
        const long long work = v4 ? total / 4 : total;
        const int blocks = (int)((work + 256 * 4 - 1) / (256 * 4) < 1 ? 1 : (work + 256 * 4 - 1) / (256 * 4) > 148 * 32 ? 148 * 32 : (work + 256 * 4 - 1) / (256 * 4));
        if (v4)
            attn_modulate_nhwc_kernel<true><<<blocks, 256, 0, st>>>(refs, attn_mul, attn_add, bias_mul, bias_add, total, C);
        else
            attn_modulate_nhwc_kernel<false><<<blocks, 256, 0, st>>>(refs, attn_mul, attn_add, bias_mul, bias_add, total, C);
        MREFSR_LAUNCH_CHECK();
        count_launches(1);
        return 0;
    }
    const bool vec = HW % 4 == 0 && al16(refs) && al16(attn_mul) && al16(attn_add);
    const int per = vec ? HW / 4 : HW;
    dim3 grid(cdiv(per, 256 * 2) < 1 ? 1 : cdiv(per, 256 * 2), planes < 32768 ? planes : 32768);
    if (vec)
        attn_modulate_kernel<true><<<grid, 256, 0, st>>>(refs, attn_mul, attn_add, bias_mul, bias_add, planes, C, HW);
    else
        attn_modulate_kernel<false><<<grid, 256, 0, st>>>(refs, attn_mul, attn_add, bias_mul, bias_add, planes, C, HW);
    MREFSR_LAUNCH_CHECK();
    count_launches(1);
    return 0;
}

int mrefsr_layout_convert(const float* src, float* dst, const float* bias, int B, int C, int HW, int to_channels_last,
                          void* stream) {
    MREFSR_CHECK(src && dst && src != dst, ERR_BAD_ARG, "layout_convert: bad pointers");
    MREFSR_CHECK(B > 0 && C > 0 && HW > 0 && C % 4 == 0, ERR_BAD_ARG, "layout_convert: needs C %% 4 == 0 (B=%d C=%d HW=%d)", B, C, HW);
    MREFSR_CHECK(al16(src) && al16(dst), ERR_BAD_ARG, "layout_convert: tensors must be 16-byte aligned");
    MREFSR_CHECK(B <= 65535 && cdiv(C, 128) <= 65535, ERR_BAD_ARG, "layout_convert: batch too large");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    ScopedTiming tm(MREFSR_K_GLUE, st);
    const dim3 grid(cdiv(HW, 32), cdiv(C, 128), B);
    if (to_channels_last)
        nchw_to_nhwc_conv_kernel<<<grid, 256, 0, st>>>(src, dst, bias, C, HW);
    else
        nhwc_to_nchw_kernel<<<grid, 256, 0, st>>>(src, dst, bias, C, HW);
    MREFSR_LAUNCH_CHECK();
    count_launches(1);
    return 0;
}

int mrefsr_layout_convert_bf16(const void* src, void* dst, int B, int C, int HW, int to_channels_last_bf16, void* stream) {
    return mrefsr_layout_convert_bf16_act(src, dst, nullptr, 1.f, B, C, HW, to_channels_last_bf16, stream);
}

int mrefsr_layout_convert_bf16_act(const void* src, void* dst, const void* gate, float slope, int B, int C, int HW,
                                   int to_channels_last_bf16, void* stream) {
    MREFSR_CHECK(src && dst && src != dst, ERR_BAD_ARG, "layout_convert_bf16: bad pointers");
    MREFSR_CHECK(!gate || (!to_channels_last_bf16 && (reinterpret_cast<uintptr_t>(gate) & 7) == 0), ERR_BAD_ARG,
                 "layout_convert_bf16: gate is an 8-byte aligned bf16 channels-last tensor of the gradient direction only");
    MREFSR_CHECK(B > 0 && C > 0 && HW > 0 && C % 4 == 0, ERR_BAD_ARG, "layout_convert_bf16: needs C %% 4 == 0 (B=%d C=%d HW=%d)", B, C, HW);
    MREFSR_CHECK((reinterpret_cast<uintptr_t>(src) & 7) == 0 && (reinterpret_cast<uintptr_t>(dst) & 7) == 0, ERR_BAD_ARG,
                 "layout_convert_bf16: tensors must be 8-byte aligned");
    MREFSR_CHECK(B <= 65535 && cdiv(C, 128) <= 65535, ERR_BAD_ARG, "layout_convert_bf16: batch too large");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    ScopedTiming tm(MREFSR_K_GLUE, st);
    const bool ct64 = C % 128 != 0 && C % 64 == 0;          // 64-channel tiles keep every lane busy at C = 64, 192, ...
    const dim3 grid(cdiv(HW, 32), cdiv(C, ct64 ? 64 : 128), B);
    if (to_channels_last_bf16) {
        if (ct64)
            nchw_f32_to_nhwc_bf16_kernel<64><<<grid, 256, 0, st>>>(static_cast<const float*>(src), static_cast<__nv_bfloat16*>(dst),
                                                                   slope, C, HW);
        else
            nchw_f32_to_nhwc_bf16_kernel<128><<<grid, 256, 0, st>>>(static_cast<const float*>(src), static_cast<__nv_bfloat16*>(dst),
                                                                    slope, C, HW);
    } else {
        if (ct64)
            nhwc_bf16_to_nchw_f32_kernel<64><<<grid, 256, 0, st>>>(static_cast<const __nv_bfloat16*>(src), static_cast<float*>(dst),
                                                                   static_cast<const __nv_bfloat16*>(gate), slope, C, HW);
        else
            nhwc_bf16_to_nchw_f32_kernel<128><<<grid, 256, 0, st>>>(static_cast<const __nv_bfloat16*>(src), static_cast<float*>(dst),
                                                                    static_cast<const __nv_bfloat16*>(gate), slope, C, HW);
    }
    MREFSR_LAUNCH_CHECK();
    count_launches(1);
    return 0;
}

int mrefsr_maxpool2x2_nhwc(const float* src, float* dst, int B, int C, int H, int W, void* stream) {
    MREFSR_CHECK(src && dst, ERR_BAD_ARG, "maxpool2x2_nhwc: null tensor");
    MREFSR_CHECK(B > 0 && C > 0 && C % 4 == 0 && H > 0 && W > 0 && H % 2 == 0 && W % 2 == 0, ERR_BAD_ARG,
                 "maxpool2x2_nhwc: needs C %% 4 == 0 and even H, W (B=%d C=%d H=%d W=%d)", B, C, H, W);
    MREFSR_CHECK(al16(src) && al16(dst), ERR_BAD_ARG, "maxpool2x2_nhwc: tensors must be 16-byte aligned");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const long long total4 = (long long)B * (H / 2) * (W / 2) * (C / 4);
    long long blocks = (total4 + 255) / 256;
    if (blocks > 148 * 32) blocks = 148 * 32;
    ScopedTiming tm(MREFSR_K_GLUE, st);
    maxpool2_nhwc_kernel<<<(int)blocks, 256, 0, st>>>(src, dst, total4, C / 4, W / 2, H / 2, W);
    MREFSR_LAUNCH_CHECK();
    count_launches(1);
    return 0;
}

}  // extern "C"

extern "C" {

int mrefsr_bias_act_train_supported(int C, int dtype) { return train_shape_ok(C, dtype) ? 1 : 0; }

int mrefsr_bias_act_train_blocks(void) { return 2 * sm_count(); }

int mrefsr_bias_act_train_forward(void* x, const float* bias, const void* residual, long long rows, int C, int dtype, int act,
                                  float slope, float scale, void* stream) {
    MREFSR_CHECK(x && rows > 0 && C > 0, ERR_BAD_ARG, "bias_act_train_forward: bad arguments");
    MREFSR_CHECK(dtype == 0 || dtype == 1, ERR_BAD_ARG, "bias_act_train_forward: dtype must be 0 (fp32) or 1 (bf16)");
    MREFSR_CHECK(act == MREFSR_ACT_NONE || act == MREFSR_ACT_LEAKY, ERR_BAD_ARG, "bias_act_train_forward: activation %d", act);
    MREFSR_CHECK(train_shape_ok(C, dtype), ERR_UNSUPPORTED, "bias_act_train_forward: C = %d not served", C);
    MREFSR_CHECK(al16(x) && (!residual || al16(residual)), ERR_BAD_ARG, "bias_act_train_forward: 16-byte alignment");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    ScopedTiming tm(MREFSR_K_GLUE, st);
    const int W = dtype ? 8 : 4;
    const long long nvec = rows * C / W;
    long long blocks = (nvec + 256 * 4 - 1) / (256 * 4);
    if (blocks > 148 * 32) blocks = 148 * 32;
    if (blocks < 1) blocks = 1;
    if (dtype)
        bias_act_train_fwd_kernel<__nv_bfloat16><<<(int)blocks, 256, 0, st>>>(
            static_cast<__nv_bfloat16*>(x), bias, static_cast<const __nv_bfloat16*>(residual), nvec, C / W, act, slope, scale);
    else
        bias_act_train_fwd_kernel<float><<<(int)blocks, 256, 0, st>>>(static_cast<float*>(x), bias,
                                                                      static_cast<const float*>(residual), nvec, C / W, act,
                                                                      slope, scale);
    MREFSR_LAUNCH_CHECK();
    count_launches(1);
    return 0;
}

int mrefsr_bias_act_train_backward(const void* grad_out, const void* y, void* grad_in, float* grad_bias, float* partial,
                                   long long rows, int C, int dtype, int act, float slope, float scale, void* stream) {
    MREFSR_CHECK(grad_out && grad_bias && partial && rows > 0 && C > 0, ERR_BAD_ARG, "bias_act_train_backward: bad arguments");
    MREFSR_CHECK(dtype == 0 || dtype == 1, ERR_BAD_ARG, "bias_act_train_backward: dtype must be 0 (fp32) or 1 (bf16)");
    MREFSR_CHECK(act == MREFSR_ACT_NONE || (act == MREFSR_ACT_LEAKY && y), ERR_BAD_ARG,
                 "bias_act_train_backward: activation %d (leaky needs the forward output)", act);
    MREFSR_CHECK(train_shape_ok(C, dtype), ERR_UNSUPPORTED, "bias_act_train_backward: C = %d not served", C);
    MREFSR_CHECK(al16(grad_out) && (!y || al16(y)) && (!grad_in || al16(grad_in)), ERR_BAD_ARG,
                 "bias_act_train_backward: 16-byte alignment");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    ScopedTiming tm(MREFSR_K_GLUE, st);
    const int W = dtype ? 8 : 4, rpb = 256 / (C / W);
    long long blocks = (rows + rpb - 1) / rpb;
    const int cap = 2 * sm_count();
    if (blocks > cap) blocks = cap;
    if (dtype)
        bias_act_train_bwd_kernel<__nv_bfloat16><<<(int)blocks, 256, 0, st>>>(
            static_cast<const __nv_bfloat16*>(grad_out), static_cast<const __nv_bfloat16*>(y),
            static_cast<__nv_bfloat16*>(grad_in), partial, rows, C, act, slope, scale);
    else
        bias_act_train_bwd_kernel<float><<<(int)blocks, 256, 0, st>>>(static_cast<const float*>(grad_out),
                                                                      static_cast<const float*>(y), static_cast<float*>(grad_in),
                                                                      partial, rows, C, act, slope, scale);
    MREFSR_LAUNCH_CHECK();
    bias_grad_sum_kernel<<<cdiv(C * 32, 256), 256, 0, st>>>(partial, grad_bias, (int)blocks, C);
    MREFSR_LAUNCH_CHECK();
    count_launches(2);
    return 0;
}

}  // extern "C"
