// 2-D neighbourhood attention core (QK^T over a clamped dilated k x k window, softmax, PV).
// One warp per (pixel, head); lanes stride the head dimension, logits live in registers and the
// softmax is a warp-shuffle reduction.  Window rule: oracle/natten_ref.py::window_start.
#pragma once
#include "cnb_common.cuh"

namespace cnb {

constexpr int NA_MAX_K = 9;
constexpr int NA_MAX_K2 = NA_MAX_K * NA_MAX_K;
constexpr int NA_MAX_DPL = 4;  // head_dim <= 128

__device__ __forceinline__ int na_window_start(int index, int length, int ksize, int dilation) {
    const int g = index % dilation;
    const int p = index / dilation;
    const int group_len = (length - g + dilation - 1) / dilation;
    int s = p - ksize / 2;
    if (s < 0) s = 0;
    if (s > group_len - ksize) s = group_len - ksize;
    return g + dilation * s;
}

template <typename T>
__global__ void __launch_bounds__(256) na2d_fwd_kernel(const T* __restrict__ qkv, T* __restrict__ out, int B, int H, int W, int heads,
                                                      int hd, int ksize, int dil, float scale) {
    const int lane = threadIdx.x & 31;
    const long warp = ((long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long nwarps = ((long)gridDim.x * blockDim.x) >> 5;
    const int C = heads * hd;
    const long total = (long)B * H * W * heads;
    const int k2 = ksize * ksize;
    for (long item = warp; item < total; item += nwarps) {
        const int head = (int)(item % heads);
        const long pix = item / heads;
        const int x = (int)(pix % W);
        const int y = (int)((pix / W) % H);
        const long img = pix / ((long)W * H);
        const int sy = na_window_start(y, H, ksize, dil);
        const int sx = na_window_start(x, W, ksize, dil);
        const T* qp = qkv + pix * 3 * C + head * hd;
        float q[NA_MAX_DPL];
#pragma unroll
        for (int j = 0; j < NA_MAX_DPL; ++j) {
            const int dd = lane + 32 * j;
            q[j] = dd < hd ? cnb_ld(qp + dd) * scale : 0.f;
        }
        float logit[NA_MAX_K2];
        float mx = -INFINITY;
        for (int n = 0; n < k2; ++n) {
            const int a = n / ksize, b = n - a * ksize;
            const long np = (img * H + (sy + a * dil)) * W + (sx + b * dil);
            const T* kp = qkv + np * 3 * C + C + head * hd;
            float part = 0.f;
#pragma unroll
            for (int j = 0; j < NA_MAX_DPL; ++j) {
                const int dd = lane + 32 * j;
                if (dd < hd) part = fmaf(q[j], cnb_ld(kp + dd), part);
            }
            const float l = cnb_warp_sum(part);
            logit[n] = l;
            mx = fmaxf(mx, l);
        }
        float den = 0.f;
        for (int n = 0; n < k2; ++n) {
            const float e = cnb_exp(logit[n] - mx);
            logit[n] = e;
            den += e;
        }
        const float inv = 1.0f / den;
        float o[NA_MAX_DPL];
#pragma unroll
        for (int j = 0; j < NA_MAX_DPL; ++j) o[j] = 0.f;
        for (int n = 0; n < k2; ++n) {
            const int a = n / ksize, b = n - a * ksize;
            const long np = (img * H + (sy + a * dil)) * W + (sx + b * dil);
            const T* vp = qkv + np * 3 * C + 2 * C + head * hd;
            const float p = logit[n] * inv;
#pragma unroll
            for (int j = 0; j < NA_MAX_DPL; ++j) {
                const int dd = lane + 32 * j;
                if (dd < hd) o[j] = fmaf(p, cnb_ld(vp + dd), o[j]);
            }
        }
        T* op = out + pix * C + head * hd;
#pragma unroll
        for (int j = 0; j < NA_MAX_DPL; ++j) {
            const int dd = lane + 32 * j;
            if (dd < hd) cnb_st(op + dd, o[j]);
        }
    }
}

// Recomputes the attention probabilities; dq is written, dk/dv are scattered with fp32 atomics into dacc.
template <typename T>
__global__ void __launch_bounds__(256) na2d_bwd_kernel(const T* __restrict__ qkv, const T* __restrict__ dout, float* __restrict__ dacc,
                                                      int B, int H, int W, int heads, int hd, int ksize, int dil, float scale) {
    const int lane = threadIdx.x & 31;
    const long warp = ((long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long nwarps = ((long)gridDim.x * blockDim.x) >> 5;
    const int C = heads * hd;
    const long total = (long)B * H * W * heads;
    const int k2 = ksize * ksize;
    for (long item = warp; item < total; item += nwarps) {
        const int head = (int)(item % heads);
        const long pix = item / heads;
        const int x = (int)(pix % W);
        const int y = (int)((pix / W) % H);
        const long img = pix / ((long)W * H);
        const int sy = na_window_start(y, H, ksize, dil);
        const int sx = na_window_start(x, W, ksize, dil);
        const T* qp = qkv + pix * 3 * C + head * hd;
        const T* dop = dout + pix * C + head * hd;
        float q[NA_MAX_DPL], dq[NA_MAX_DPL], go[NA_MAX_DPL];
#pragma unroll
        for (int j = 0; j < NA_MAX_DPL; ++j) {
            const int dd = lane + 32 * j;
            q[j] = dd < hd ? cnb_ld(qp + dd) * scale : 0.f;
            go[j] = dd < hd ? cnb_ld(dop + dd) : 0.f;
            dq[j] = 0.f;
        }
        float prob[NA_MAX_K2], dprob[NA_MAX_K2];
        float mx = -INFINITY;
        for (int n = 0; n < k2; ++n) {
            const int a = n / ksize, b = n - a * ksize;
            const long np = (img * H + (sy + a * dil)) * W + (sx + b * dil);
            const T* kp = qkv + np * 3 * C + C + head * hd;
            const T* vp = kp + C;
            float part = 0.f, dpart = 0.f;
#pragma unroll
            for (int j = 0; j < NA_MAX_DPL; ++j) {
                const int dd = lane + 32 * j;
                if (dd < hd) {
                    part = fmaf(q[j], cnb_ld(kp + dd), part);
                    dpart = fmaf(go[j], cnb_ld(vp + dd), dpart);
                }
            }
            const float l = cnb_warp_sum(part);
            prob[n] = l;
            dprob[n] = cnb_warp_sum(dpart);
            mx = fmaxf(mx, l);
        }
        float den = 0.f;
        for (int n = 0; n < k2; ++n) {
            const float e = cnb_exp(prob[n] - mx);
            prob[n] = e;
            den += e;
        }
        const float inv = 1.0f / den;
        float dot = 0.f;
        for (int n = 0; n < k2; ++n) {
            prob[n] *= inv;
            dot = fmaf(prob[n], dprob[n], dot);
        }
        for (int n = 0; n < k2; ++n) {
            const int a = n / ksize, b = n - a * ksize;
            const long np = (img * H + (sy + a * dil)) * W + (sx + b * dil);
            const T* kp = qkv + np * 3 * C + C + head * hd;
            float* dkp = dacc + np * 3 * C + C + head * hd;
            float* dvp = dkp + C;
            const float p = prob[n];
            const float ds = p * (dprob[n] - dot);
#pragma unroll
            for (int j = 0; j < NA_MAX_DPL; ++j) {
                const int dd = lane + 32 * j;
                if (dd < hd) {
                    dq[j] = fmaf(ds, cnb_ld(kp + dd), dq[j]);
                    atomicAdd(dkp + dd, ds * q[j]);
                    atomicAdd(dvp + dd, p * go[j]);
                }
            }
        }
        float* dqp = dacc + pix * 3 * C + head * hd;
#pragma unroll
        for (int j = 0; j < NA_MAX_DPL; ++j) {
            const int dd = lane + 32 * j;
            if (dd < hd) dqp[dd] = dq[j] * scale;
        }
    }
}

template <typename T>
__global__ void __launch_bounds__(256) cast_from_f32_kernel(const float* __restrict__ src, T* __restrict__ dst, long n) {
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) cnb_st(dst + i, src[i]);
}

}  // namespace cnb
