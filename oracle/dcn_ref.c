/*
 * oracle/dcn_ref.c -- plain-C CPU restatement of DCNv2 forward/backward.
 *
 * TEST INFRASTRUCTURE ONLY (see oracle/__init__.py): used by tests/ as a second
 * opinion next to oracle/dcn.py, and by bench.py's cpu_baseline / --impl reference
 * legs as the multi-threaded CPU baseline.  Never linked into the product library.
 *
 * Follows the algorithm of the reference extension
 *   basicsr/ops/dcn/src/deform_conv_cuda_kernel.cu  :468-497 bilinear sampling,
 *       :571-633 modulated im2col, :635-693 col2im, :695-767 col2im_coord
 *   basicsr/ops/dcn/src/deform_conv_cuda.cpp        :490-569 forward, :571-685 backward
 * i.e. per sample: columns = deformable-im2col(input, offset, mask);
 *                  out[g] = W[g] . columns[g] (+ bias);
 * backward: gcol = W^T . gout; grad_offset/grad_mask from gcol and the image,
 * grad_input by bilinear scatter of gcol*mask, grad_weight += gout . columns^T,
 * grad_bias += sum gout.  Loop structure and naming are this file's own.
 *
 * Build: gcc -O3 -fopenmp -shared -fPIC (oracle/build.py).
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
    int B, C, H, W, Co, kh, kw, sh, sw, ph, pw, dh, dw, G, DG, Ho, Wo;
} dcn_shape;

static dcn_shape mk(int B, int C, int H, int W, int Co, int kh, int kw, int sh, int sw, int ph, int pw,
                    int dh, int dw, int G, int DG) {
    dcn_shape s = {B, C, H, W, Co, kh, kw, sh, sw, ph, pw, dh, dw, G, DG, 0, 0};
    s.Ho = (H + 2 * ph - (dh * (kh - 1) + 1)) / sh + 1;
    s.Wo = (W + 2 * pw - (dw * (kw - 1) + 1)) / sw + 1;
    return s;
}

/* value of plane at fractional (y, x); corners outside the plane read as 0 (.cu:468-497) */
static float sample(const float *plane, int H, int W, float y, float x) {
    int y0 = (int)floorf(y), x0 = (int)floorf(x);
    int y1 = y0 + 1, x1 = x0 + 1;
    float ly = y - y0, lx = x - x0, hy = 1.f - ly, hx = 1.f - lx;
    float a = (y0 >= 0 && x0 >= 0) ? plane[y0 * W + x0] : 0.f;
    float b = (y0 >= 0 && x1 <= W - 1) ? plane[y0 * W + x1] : 0.f;
    float c = (y1 <= H - 1 && x0 >= 0) ? plane[y1 * W + x0] : 0.f;
    float d = (y1 <= H - 1 && x1 <= W - 1) ? plane[y1 * W + x1] : 0.f;
    return hy * hx * a + hy * lx * b + ly * hx * c + ly * lx * d;
}

/* sampling position of tap (i, j) of deform group g at output (oy, ox) of sample b */
static void tap_pos(const dcn_shape *s, const float *offset, const float *mask, int b, int g, int i, int j,
                    int oy, int ox, float *y, float *x, float *m) {
    int K = s->kh * s->kw, k = i * s->kw + j, P = s->Ho * s->Wo, p = oy * s->Wo + ox;
    const float *ob = offset + ((size_t)(b * s->DG + g) * 2 * K) * P;
    const float *mb = mask + ((size_t)(b * s->DG + g) * K) * P;
    *y = (float)(oy * s->sh - s->ph + i * s->dh) + ob[(size_t)(2 * k) * P + p];
    *x = (float)(ox * s->sw - s->pw + j * s->dw) + ob[(size_t)(2 * k + 1) * P + p];
    *m = mb[(size_t)k * P + p];
}

/* columns[(c*K + k), p] for one sample (.cu:571-633) */
static void im2col(const dcn_shape *s, const float *x, const float *offset, const float *mask, int b, float *col) {
    int K = s->kh * s->kw, P = s->Ho * s->Wo, cg = s->C / s->DG;
#pragma omp parallel for schedule(static)
    for (int c = 0; c < s->C; ++c) {
        const float *plane = x + ((size_t)b * s->C + c) * s->H * s->W;
        int g = c / cg;
        for (int i = 0; i < s->kh; ++i)
            for (int j = 0; j < s->kw; ++j) {
                float *dst = col + ((size_t)c * K + i * s->kw + j) * P;
                for (int oy = 0; oy < s->Ho; ++oy)
                    for (int ox = 0; ox < s->Wo; ++ox) {
                        float y, xx, m, v = 0.f;
                        tap_pos(s, offset, mask, b, g, i, j, oy, ox, &y, &xx, &m);
                        if (y > -1 && xx > -1 && y < s->H && xx < s->W) v = sample(plane, s->H, s->W, y, xx);
                        dst[oy * s->Wo + ox] = v * m;
                    }
            }
    }
}

void dcn_ref_forward(const float *x, const float *offset, const float *mask, const float *weight,
                     const float *bias, float *out, int B, int C, int H, int W, int Co, int kh, int kw, int sh,
                     int sw, int ph, int pw, int dh, int dw, int G, int DG, int with_bias) {
    dcn_shape s = mk(B, C, H, W, Co, kh, kw, sh, sw, ph, pw, dh, dw, G, DG);
    int K = kh * kw, P = s.Ho * s.Wo, cpg = C / G, opg = Co / G, KK = cpg * K;
    float *col = (float *)malloc((size_t)C * K * P * sizeof(float));
    for (int b = 0; b < B; ++b) {
        im2col(&s, x, offset, mask, b, col);
#pragma omp parallel for schedule(static)
        for (int o = 0; o < Co; ++o) {
            int g = o / opg;
            float *dst = out + ((size_t)b * Co + o) * P;
            const float *wrow = weight + (size_t)o * KK;
            float b0 = (with_bias && bias) ? bias[o] : 0.f;
            for (int p = 0; p < P; ++p) dst[p] = 0.f;
            for (int q = 0; q < KK; ++q) {
                float wv = wrow[q];
                const float *src = col + ((size_t)g * KK + q) * P;
                for (int p = 0; p < P; ++p) dst[p] += wv * src[p];
            }
            for (int p = 0; p < P; ++p) dst[p] += b0;
        }
    }
    free(col);
}

void dcn_ref_backward(const float *x, const float *offset, const float *mask, const float *weight,
                      const float *gout, float *gx, float *goff, float *gmask, float *gw, float *gb,
                      const float *unused, int B, int C, int H, int W, int Co, int kh, int kw, int sh, int sw,
                      int ph, int pw, int dh, int dw, int G, int DG, int with_bias) {
    (void)unused;
    dcn_shape s = mk(B, C, H, W, Co, kh, kw, sh, sw, ph, pw, dh, dw, G, DG);
    int K = kh * kw, P = s.Ho * s.Wo, cpg = C / G, opg = Co / G, KK = cpg * K, cg = C / DG;
    float *col = (float *)malloc((size_t)C * K * P * sizeof(float));
    float *gcol = (float *)malloc((size_t)C * K * P * sizeof(float));
    for (int b = 0; b < B; ++b) {
        /* gcol = W^T . gout   (.cpp:623-626) */
#pragma omp parallel for schedule(static)
        for (int r = 0; r < C * K; ++r) {
            int g = r / KK, q = r % KK;
            float *dst = gcol + (size_t)r * P;
            for (int p = 0; p < P; ++p) dst[p] = 0.f;
            for (int oo = 0; oo < opg; ++oo) {
                int o = g * opg + oo;
                float wv = weight[(size_t)o * KK + q];
                const float *src = gout + ((size_t)b * Co + o) * P;
                for (int p = 0; p < P; ++p) dst[p] += wv * src[p];
            }
        }
        /* grad_offset / grad_mask (.cu:695-767) */
#pragma omp parallel for schedule(static)
        for (int gk = 0; gk < DG * K; ++gk) {
            int g = gk / K, k = gk % K, i = k / kw, j = k % kw;
            for (int oy = 0; oy < s.Ho; ++oy)
                for (int ox = 0; ox < s.Wo; ++ox) {
                    int p = oy * s.Wo + ox;
                    float y, xx, m;
                    tap_pos(&s, offset, mask, b, g, i, j, oy, ox, &y, &xx, &m);
                    float dy = 0.f, dx = 0.f, dm = 0.f;
                    if (y > -1 && xx > -1 && y < H && xx < W) {
                        int y0 = (int)floorf(y), x0 = (int)floorf(xx), y1 = y0 + 1, x1 = x0 + 1;
                        float ly = y - y0, lx = xx - x0;
                        for (int cc = 0; cc < cg; ++cc) {
                            int c = g * cg + cc;
                            const float *pl = x + ((size_t)b * C + c) * H * W;
                            float a = (y0 >= 0 && x0 >= 0) ? pl[y0 * W + x0] : 0.f;
                            float bq = (y0 >= 0 && x1 <= W - 1) ? pl[y0 * W + x1] : 0.f;
                            float cq = (y1 <= H - 1 && x0 >= 0) ? pl[y1 * W + x0] : 0.f;
                            float d = (y1 <= H - 1 && x1 <= W - 1) ? pl[y1 * W + x1] : 0.f;
                            float gc = gcol[((size_t)c * K + k) * P + p];
                            float v = (1 - ly) * (1 - lx) * a + (1 - ly) * lx * bq + ly * (1 - lx) * cq + ly * lx * d;
                            dm += gc * v;
                            dy += gc * m * ((1 - lx) * (cq - a) + lx * (d - bq));
                            dx += gc * m * ((1 - ly) * (bq - a) + ly * (d - cq));
                        }
                    }
                    goff[(((size_t)(b * DG + g) * 2 * K) + 2 * k) * P + p] = dy;
                    goff[(((size_t)(b * DG + g) * 2 * K) + 2 * k + 1) * P + p] = dx;
                    gmask[(((size_t)(b * DG + g) * K) + k) * P + p] = dm;
                }
        }
        /* grad_input: bilinear scatter of gcol*mask (.cu:635-693); parallel over channels => no races */
#pragma omp parallel for schedule(static)
        for (int c = 0; c < C; ++c) {
            int g = c / cg;
            float *pl = gx + ((size_t)b * C + c) * H * W;
            for (int k = 0; k < K; ++k) {
                int i = k / kw, j = k % kw;
                for (int oy = 0; oy < s.Ho; ++oy)
                    for (int ox = 0; ox < s.Wo; ++ox) {
                        int p = oy * s.Wo + ox;
                        float y, xx, m;
                        tap_pos(&s, offset, mask, b, g, i, j, oy, ox, &y, &xx, &m);
                        if (!(y > -1 && xx > -1 && y < H && xx < W)) continue;
                        float t = gcol[((size_t)c * K + k) * P + p] * m;
                        int y0 = (int)floorf(y), x0 = (int)floorf(xx), y1 = y0 + 1, x1 = x0 + 1;
                        float ly = y - y0, lx = xx - x0;
                        if (y0 >= 0 && x0 >= 0) pl[y0 * W + x0] += (1 - ly) * (1 - lx) * t;
                        if (y0 >= 0 && x1 <= W - 1) pl[y0 * W + x1] += (1 - ly) * lx * t;
                        if (y1 <= H - 1 && x0 >= 0) pl[y1 * W + x0] += ly * (1 - lx) * t;
                        if (y1 <= H - 1 && x1 <= W - 1) pl[y1 * W + x1] += ly * lx * t;
                    }
            }
        }
        /* grad_weight += gout . columns^T ; grad_bias += sum gout (.cpp:647-671) */
        im2col(&s, x, offset, mask, b, col);
#pragma omp parallel for schedule(static)
        for (int o = 0; o < Co; ++o) {
            int g = o / opg;
            const float *go = gout + ((size_t)b * Co + o) * P;
            for (int q = 0; q < KK; ++q) {
                const float *src = col + ((size_t)g * KK + q) * P;
                float acc = 0.f;
                for (int p = 0; p < P; ++p) acc += go[p] * src[p];
                gw[(size_t)o * KK + q] += acc;
            }
            if (with_bias) {
                float acc = 0.f;
                for (int p = 0; p < P; ++p) acc += go[p];
                gb[o] += acc;
            }
        }
    }
    free(col);
    free(gcol);
}
