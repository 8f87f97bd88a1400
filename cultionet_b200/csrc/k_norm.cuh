// BatchNorm(+SiLU), LayerNorm and n-ary add over pixel-major [P][L] matrices.  HBM-bound kernels:
// each pass streams its operands once; per-channel reductions stay in registers -> shared memory ->
// one global atomic per (CTA, channel).
#pragma once
#include "cnb_common.cuh"

namespace cnb {

constexpr int BN_SMEM_C = 512;  // channels reduced through shared memory; above this, global atomics directly

// Thread `gid` walks elements gid, gid+stride, ... with stride a multiple of L, so it always sits on one column.
template <typename T>
__global__ void __launch_bounds__(256) bn_stats_kernel(const T* __restrict__ x, long total, int L, int C, int ch_div, long stride,
                                                      float* __restrict__ sums) {
    CNB_PDL_SYNC();
    __shared__ float sh[2 * BN_SMEM_C];
    const bool use_sh = C <= BN_SMEM_C;
    if (use_sh) {
        for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) sh[i] = 0.f;
        __syncthreads();
    }
    const long gid = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid < stride && (int)(gid % L) / ch_div < C) {
        const int col = (int)(gid % L);
        const int ch = col / ch_div;  // columns with ch >= C are row padding and take no part
        float s = 0.f, s2 = 0.f;
        for (long i = gid; i < total; i += stride) {
            const float v = cnb_ld(x + i);
            s += v;
            s2 = fmaf(v, v, s2);
        }
        if (use_sh) {
            atomicAdd(&sh[ch], s);
            atomicAdd(&sh[C + ch], s2);
        } else {
            atomicAdd(&sums[ch], s);
            atomicAdd(&sums[C + ch], s2);
        }
    }
    if (use_sh) {
        __syncthreads();
        for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) {
            const float v = sh[i];
            if (v != 0.f) atomicAdd(&sums[i], v);
        }
    }
}

__global__ void bn_finalize_kernel(const float* __restrict__ sums, long count, int C, const float* __restrict__ gamma,
                                   const float* __restrict__ beta, float eps, float momentum, float* running_mean, float* running_var,
                                   float* save_mean, float* save_rstd, float* scale, float* shift) {
    CNB_PDL_SYNC();
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    float mean, var;
    if (sums) {
        const float inv = 1.0f / (float)count;
        mean = sums[c] * inv;
        var = fmaxf(sums[C + c] * inv - mean * mean, 0.f);
        if (running_mean) {
            const float unbiased = count > 1 ? var * ((float)count / (float)(count - 1)) : var;
            running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * mean;
            running_var[c] = (1.f - momentum) * running_var[c] + momentum * unbiased;
        }
    } else {
        mean = running_mean[c];
        var = running_var[c];
    }
    const float rstd = rsqrtf(var + eps);
    const float g = gamma ? gamma[c] : 1.f;
    const float b = beta ? beta[c] : 0.f;
    if (save_mean) save_mean[c] = mean;
    if (save_rstd) save_rstd[c] = rstd;
    scale[c] = g * rstd;
    shift[c] = b - mean * g * rstd;
}

template <typename T>
__global__ void __launch_bounds__(256) bn_act_fwd_kernel(const T* __restrict__ x, const float* __restrict__ scale,
                                                        const float* __restrict__ shift, const T* __restrict__ residual,
                                                        T* __restrict__ y, long total, int L, int C, int ch_div, int act) {
    CNB_PDL_SYNC();
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        const int col = (int)(i % L);
        const int ch = col / ch_div;
        if (ch >= C) {  // row padding stays zero
            cnb_st(y + i, 0.f);
            continue;
        }
        float z = fmaf(cnb_ld(x + i), scale[ch], shift[ch]);
        z = cnb_act_t<float>(z, act);
        if (residual) z += cnb_ld(residual + i);
        cnb_st(y + i, z);
    }
}

template <typename T>
__global__ void __launch_bounds__(256) bn_act_bwd_reduce_kernel(const T* __restrict__ x, const T* __restrict__ dy,
                                                               const float* __restrict__ mean, const float* __restrict__ rstd,
                                                               const float* __restrict__ gamma, const float* __restrict__ beta,
                                                               long total, int L, int C, int ch_div, int act, long stride,
                                                               float* __restrict__ dsums) {
    CNB_PDL_SYNC();
    __shared__ float sh[2 * BN_SMEM_C];
    const bool use_sh = C <= BN_SMEM_C;
    if (use_sh) {
        for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) sh[i] = 0.f;
        __syncthreads();
    }
    const long gid = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid < stride && (int)(gid % L) / ch_div < C) {
        const int col = (int)(gid % L);
        const int ch = col / ch_div;
        const float mu = mean[ch], rs = rstd[ch];
        const float g = gamma ? gamma[ch] : 1.f, b = beta ? beta[ch] : 0.f;
        float s = 0.f, sx = 0.f;
        for (long i = gid; i < total; i += stride) {
            const float xh = (cnb_ld(x + i) - mu) * rs;
            float dz = cnb_ld(dy + i);
            if (act) dz *= cnb_act_grad_t<float>(fmaf(xh, g, b), act);
            s += dz;
            sx = fmaf(dz, xh, sx);
        }
        if (use_sh) {
            atomicAdd(&sh[ch], s);
            atomicAdd(&sh[C + ch], sx);
        } else {
            atomicAdd(&dsums[ch], s);
            atomicAdd(&dsums[C + ch], sx);
        }
    }
    if (use_sh) {
        __syncthreads();
        for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) {
            const float v = sh[i];
            if (v != 0.f) atomicAdd(&dsums[i], v);
        }
    }
}

template <typename T>
__global__ void __launch_bounds__(256) bn_act_bwd_apply_kernel(const T* __restrict__ x, const T* __restrict__ dy,
                                                              const float* __restrict__ mean, const float* __restrict__ rstd,
                                                              const float* __restrict__ gamma, const float* __restrict__ beta,
                                                              const float* __restrict__ dsums, float inv_count, T* __restrict__ dx,
                                                              long total, int L, int C, int ch_div, int act, int train_stats) {
    CNB_PDL_SYNC();
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        const int col = (int)(i % L);
        const int ch = col / ch_div;
        if (ch >= C) {  // row padding: zero gradient
            cnb_st(dx + i, 0.f);
            continue;
        }
        const float rs = rstd[ch];
        const float g = gamma ? gamma[ch] : 1.f, b = beta ? beta[ch] : 0.f;
        const float xh = (cnb_ld(x + i) - mean[ch]) * rs;
        float dz = cnb_ld(dy + i);
        if (act) dz *= cnb_act_grad_t<float>(fmaf(xh, g, b), act);
        float r;
        if (train_stats)
            r = g * rs * (dz - dsums[ch] * inv_count - xh * dsums[C + ch] * inv_count);
        else
            r = g * rs * dz;
        cnb_st(dx + i, r);
    }
}

template <typename T>
__global__ void __launch_bounds__(256) add_n_kernel(const T* __restrict__ a, const T* __restrict__ b, const T* __restrict__ c,
                                                   const T* __restrict__ d, T* __restrict__ out, long n) {
    CNB_PDL_SYNC();
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
        float v = cnb_ld(a + i) + cnb_ld(b + i);
        if (c) v += cnb_ld(c + i);
        if (d) v += cnb_ld(d + i);
        cnb_st(out + i, v);
    }
}

// ---------------------------------------------------------------------------------------------
// LayerNorm over channels: one warp per pixel row, lanes stride the channels
// ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) layernorm_fwd_kernel(const T* __restrict__ x, const float* __restrict__ gamma,
                                                           const float* __restrict__ beta, float eps, T* __restrict__ y,
                                                           float* __restrict__ save_mean, float* __restrict__ save_rstd, long P, int C) {
    CNB_PDL_SYNC();
    const int lane = threadIdx.x & 31;
    const long warp = ((long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long nwarps = ((long)gridDim.x * blockDim.x) >> 5;
    for (long p = warp; p < P; p += nwarps) {
        const T* xr = x + p * C;
        float s = 0.f;
        for (int c = lane; c < C; c += 32) s += cnb_ld(xr + c);
        const float mean = cnb_warp_sum(s) / (float)C;
        float v = 0.f;
        for (int c = lane; c < C; c += 32) {
            const float dlt = cnb_ld(xr + c) - mean;
            v = fmaf(dlt, dlt, v);
        }
        const float rstd = rsqrtf(cnb_warp_sum(v) / (float)C + eps);
        for (int c = lane; c < C; c += 32) cnb_st(y + p * C + c, (cnb_ld(xr + c) - mean) * rstd * gamma[c] + beta[c]);
        if (lane == 0) {
            save_mean[p] = mean;
            save_rstd[p] = rstd;
        }
    }
}

constexpr int LN_MAX_CPL = 32;  // channels per lane held in registers for dgamma/dbeta -> C <= 1024

template <typename T>
__global__ void __launch_bounds__(256) layernorm_bwd_kernel(const T* __restrict__ x, const T* __restrict__ dy,
                                                           const float* __restrict__ gamma, const float* __restrict__ save_mean,
                                                           const float* __restrict__ save_rstd, T* __restrict__ dx,
                                                           float* __restrict__ dgamma, float* __restrict__ dbeta, long P, int C) {
    CNB_PDL_SYNC();
    const int lane = threadIdx.x & 31;
    const long warp = ((long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long nwarps = ((long)gridDim.x * blockDim.x) >> 5;
    float dg[LN_MAX_CPL], db[LN_MAX_CPL];
#pragma unroll
    for (int j = 0; j < LN_MAX_CPL; ++j) {
        dg[j] = 0.f;
        db[j] = 0.f;
    }
    for (long p = warp; p < P; p += nwarps) {
        const T* xr = x + p * C;
        const T* dyr = dy + p * C;
        const float mean = save_mean[p], rstd = save_rstd[p];
        float s1 = 0.f, s2 = 0.f;
        for (int c = lane; c < C; c += 32) {
            const float xh = (cnb_ld(xr + c) - mean) * rstd;
            const float g = cnb_ld(dyr + c) * gamma[c];
            s1 += g;
            s2 = fmaf(g, xh, s2);
        }
        s1 = cnb_warp_sum(s1) / (float)C;
        s2 = cnb_warp_sum(s2) / (float)C;
#pragma unroll
        for (int j = 0; j < LN_MAX_CPL; ++j) {
            const int c = lane + 32 * j;
            if (c < C) {
                const float xh = (cnb_ld(xr + c) - mean) * rstd;
                const float d = cnb_ld(dyr + c);
                cnb_st(dx + p * C + c, rstd * (d * gamma[c] - s1 - xh * s2));
                dg[j] = fmaf(d, xh, dg[j]);
                db[j] += d;
            }
        }
    }
#pragma unroll
    for (int j = 0; j < LN_MAX_CPL; ++j) {
        const int c = lane + 32 * j;
        if (c < C) {
            atomicAdd(dgamma + c, dg[j]);
            atomicAdd(dbeta + c, db[j]);
        }
    }
}

}  // namespace cnb
