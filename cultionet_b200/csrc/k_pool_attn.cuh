// Bandwidth-bound kernels of the optional ResUNet-a block variants over pixel-major [B,H,W,C] activations:
//   * adaptive max pooling to (H/2, W/2)                      reference nn/modules/convolution.py:499-503 (pool_by_max=True)
//   * spatial-channel attention (CBAM-style)                   reference nn/modules/attention.py:12-125, convolution.py:355-360,:392-393
//   * SiLU as a stand-alone operator (the channel MLP of the attention block)
//   * Dropout2d / Dropout with a counter-based generator       reference convolution.py:487 (Dropout2d), natten proj_drop
//
// Every kernel is templated on the element type T (float / bf16) and a vector width VEC (1, or 16 bytes' worth of elements when the
// channel count allows whole vectors): a thread moves VEC consecutive channels of one pixel per step, so warps read whole lines.
// All small tensors of the attention block (pooled statistics, logits, their gradients) are fp32.
#pragma once
#include "cnb_common.cuh"

namespace cnb {

template <typename T, int VEC>
__device__ __forceinline__ void pa_ld(const T* p, float* v) {
    if constexpr (VEC == 1)
        v[0] = cnb_ld(p);
    else
        cnb_ldv(p, v);
}
template <typename T, int VEC>
__device__ __forceinline__ void pa_st(T* p, const float* v) {
    if constexpr (VEC == 1)
        cnb_st(p, v[0]);
    else
        cnb_stv(p, v);
}

// ---------------------------------------------------------------------------------------------------------------------
// adaptive max pooling (torch.nn.functional.adaptive_max_pool2d): output o covers input [floor(o*in/out), ceil((o+1)*in/out));
// the first maximum in (row, column) scan order wins, its position inside the window is kept as one byte (row*16 + column)
// ---------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ int ap_start(int o, int in, int out) { return (int)(((long)o * in) / out); }
__device__ __forceinline__ int ap_end(int o, int in, int out) { return (int)(((long)(o + 1) * in + out - 1) / out); }

template <typename T, int VEC>
__global__ void __launch_bounds__(256) adaptive_maxpool_fwd_kernel(const T* __restrict__ x, T* __restrict__ y, uint8_t* __restrict__ idx,
                                                                  int B, int Hin, int Win, int Hout, int Wout, int C) {
    CNB_PDL_SYNC();
    const int CV = C / VEC;
    const long total = (long)B * Hout * Wout * CV;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        const int cv = (int)(i % CV);
        long p = i / CV;
        const int ow = (int)(p % Wout);
        p /= Wout;
        const int oh = (int)(p % Hout);
        const int b = (int)(p / Hout);
        const int hs = ap_start(oh, Hin, Hout), he = ap_end(oh, Hin, Hout);
        const int ws = ap_start(ow, Win, Wout), we = ap_end(ow, Win, Wout);
        float best[VEC];
        int code[VEC];
#pragma unroll
        for (int j = 0; j < VEC; ++j) best[j] = -INFINITY, code[j] = 0;
        for (int ih = hs; ih < he; ++ih)
            for (int iw = ws; iw < we; ++iw) {
                float v[VEC];
                pa_ld<T, VEC>(x + (((long)b * Hin + ih) * Win + iw) * C + cv * VEC, v);
#pragma unroll
                for (int j = 0; j < VEC; ++j)
                    if (v[j] > best[j] || v[j] != v[j]) best[j] = v[j], code[j] = (ih - hs) * 16 + (iw - ws);
            }
        const long o = (((long)b * Hout + oh) * Wout + ow) * C + cv * VEC;
        pa_st<T, VEC>(y + o, best);
#pragma unroll
        for (int j = 0; j < VEC; ++j) idx[o + j] = (uint8_t)code[j];
    }
}

// gather form of the backward: an input pixel sums the gradients of the (at most 2 x 2) windows that contain it and chose it
template <typename T, int VEC>
__global__ void __launch_bounds__(256) adaptive_maxpool_bwd_kernel(const T* __restrict__ dy, const uint8_t* __restrict__ idx,
                                                                  T* __restrict__ dx, int B, int Hin, int Win, int Hout, int Wout, int C) {
    CNB_PDL_SYNC();
    const int CV = C / VEC;
    const long total = (long)B * Hin * Win * CV;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        const int cv = (int)(i % CV);
        long p = i / CV;
        const int iw = (int)(p % Win);
        p /= Win;
        const int ih = (int)(p % Hin);
        const int b = (int)(p / Hin);
        float acc[VEC];
#pragma unroll
        for (int j = 0; j < VEC; ++j) acc[j] = 0.f;
        const int ohc = (int)(((long)ih * Hout) / Hin), owc = (int)(((long)iw * Wout) / Win);
        for (int oh = ohc - 1; oh <= ohc + 1; ++oh) {
            if (oh < 0 || oh >= Hout) continue;
            const int hs = ap_start(oh, Hin, Hout);
            if (ih < hs || ih >= ap_end(oh, Hin, Hout)) continue;
            for (int ow = owc - 1; ow <= owc + 1; ++ow) {
                if (ow < 0 || ow >= Wout) continue;
                const int ws = ap_start(ow, Win, Wout);
                if (iw < ws || iw >= ap_end(ow, Win, Wout)) continue;
                const int want = (ih - hs) * 16 + (iw - ws);
                const long o = (((long)b * Hout + oh) * Wout + ow) * C + cv * VEC;
                float g[VEC];
                pa_ld<T, VEC>(dy + o, g);
#pragma unroll
                for (int j = 0; j < VEC; ++j)
                    if ((int)idx[o + j] == want) acc[j] += g[j];
            }
        }
        pa_st<T, VEC>(dx + (((long)b * Hin + ih) * Win + iw) * C + cv * VEC, acc);
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// stand-alone activation (fp32 or bf16); act = a CNB_ACT_* code
// ---------------------------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) act_fwd_kernel(const T* __restrict__ x, T* __restrict__ y, long n, int act) {
    CNB_PDL_SYNC();
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x)
        cnb_st(y + i, cnb_act_t<T>(cnb_ld(x + i), act));
}
template <typename T>
__global__ void __launch_bounds__(256) act_bwd_kernel(const T* __restrict__ x, const T* __restrict__ dy, T* __restrict__ dx, long n, int act) {
    CNB_PDL_SYNC();
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x)
        cnb_st(dx + i, cnb_ld(dy + i) * cnb_act_grad_t<T>(cnb_ld(x + i), act));
}

// ---------------------------------------------------------------------------------------------------------------------
// spatial-channel attention, pooling side
//   spatial branch (attention.py:78-86): per pixel mean and max over the channels (einops reduce -> torch.amax: a tie shares the
//     gradient evenly, so the number of maxima is kept);
//   channel branch (attention.py:17-18, :56-57): per (sample, channel) mean and max over the pixels (AdaptiveMaxPool2d(1): the first
//     maximum in scan order receives the gradient, its pixel index is kept).
// ---------------------------------------------------------------------------------------------------------------------
// one warp per pixel; sp[p] = (mean, max), ties[p] = number of channels equal to the max
template <typename T, int VEC>
__global__ void __launch_bounds__(256) sca_spatial_pool_kernel(const T* __restrict__ x, float* __restrict__ sp, float* __restrict__ ties,
                                                              long P, int C) {
    CNB_PDL_SYNC();
    const int lane = threadIdx.x & 31;
    const long warp = ((long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long nwarps = ((long)gridDim.x * blockDim.x) >> 5;
    const int CV = C / VEC;
    for (long p = warp; p < P; p += nwarps) {
        const T* xr = x + p * C;
        float s = 0.f, m = -INFINITY;
        for (int g = lane; g < CV; g += 32) {
            float v[VEC];
            pa_ld<T, VEC>(xr + g * VEC, v);
#pragma unroll
            for (int j = 0; j < VEC; ++j) s += v[j], m = fmaxf(m, v[j]);
        }
        s = cnb_warp_sum(s);
        m = cnb_warp_max(m);
        float n = 0.f;
        for (int g = lane; g < CV; g += 32) {
            float v[VEC];
            pa_ld<T, VEC>(xr + g * VEC, v);
#pragma unroll
            for (int j = 0; j < VEC; ++j) n += v[j] == m ? 1.f : 0.f;
        }
        n = cnb_warp_sum(n);
        if (lane == 0) {
            sp[2 * p] = s / (float)C;
            sp[2 * p + 1] = m;
            ties[p] = n;
        }
    }
}

constexpr int SCA_THREADS = 256;

// grid (S slices, B samples, column tiles): thread (row r, column group g) walks the pixels r, r+R, ... of its slice and keeps the
// sum / max / first arg-max of its VEC channels in registers; the rows of the CTA are then merged through shared memory.
template <typename T, int VEC>
__global__ void __launch_bounds__(SCA_THREADS) sca_channel_pool_partial_kernel(const T* __restrict__ x, float* __restrict__ part_sum,
                                                                              float* __restrict__ part_max, int* __restrict__ part_arg,
                                                                              int HW, int C, int S) {
    CNB_PDL_SYNC();
    __shared__ float sh_sum[SCA_THREADS * VEC];
    __shared__ float sh_max[SCA_THREADS * VEC];
    __shared__ int sh_arg[SCA_THREADS * VEC];
    const int CV = C / VEC;
    const int cols = CV < SCA_THREADS ? CV : SCA_THREADS;
    const int R = SCA_THREADS / cols;
    const int r = threadIdx.x / cols, gl = threadIdx.x % cols;
    const int g = blockIdx.z * cols + gl;
    const bool active = r < R && g < CV;
    const int s = blockIdx.x, b = blockIdx.y;
    const int chunk = (HW + S - 1) / S;
    const int p0 = s * chunk, p1 = p0 + chunk < HW ? p0 + chunk : HW;
    float sum[VEC], mx[VEC];
    int arg[VEC];
#pragma unroll
    for (int j = 0; j < VEC; ++j) sum[j] = 0.f, mx[j] = -INFINITY, arg[j] = p0;
    if (active)
        for (int p = p0 + r; p < p1; p += R) {
            float v[VEC];
            pa_ld<T, VEC>(x + ((long)b * HW + p) * C + g * VEC, v);
#pragma unroll
            for (int j = 0; j < VEC; ++j) {
                sum[j] += v[j];
                if (v[j] > mx[j] || v[j] != v[j]) mx[j] = v[j], arg[j] = p;
            }
        }
#pragma unroll
    for (int j = 0; j < VEC; ++j) {
        sh_sum[threadIdx.x * VEC + j] = sum[j];
        sh_max[threadIdx.x * VEC + j] = mx[j];
        sh_arg[threadIdx.x * VEC + j] = arg[j];
    }
    __syncthreads();
    if (active && r == 0) {
        for (int rr = 1; rr < R; ++rr) {
            const int t = rr * cols + gl;
#pragma unroll
            for (int j = 0; j < VEC; ++j) {
                sum[j] += sh_sum[t * VEC + j];
                const float m2 = sh_max[t * VEC + j];
                const int a2 = sh_arg[t * VEC + j];
                if (m2 > mx[j] || (m2 == mx[j] && a2 < arg[j])) mx[j] = m2, arg[j] = a2;
            }
        }
        const long o = ((long)b * S + s) * C + g * VEC;
#pragma unroll
        for (int j = 0; j < VEC; ++j) part_sum[o + j] = sum[j], part_max[o + j] = mx[j], part_arg[o + j] = arg[j];
    }
}

// ch[b] = (mean[C], max[C]) and arg[b][C] from the S partials
__global__ void __launch_bounds__(256) sca_channel_pool_final_kernel(const float* __restrict__ part_sum, const float* __restrict__ part_max,
                                                                    const int* __restrict__ part_arg, float* __restrict__ ch_avg,
                                                                    float* __restrict__ ch_max, int* __restrict__ ch_arg, int B, int C,
                                                                    int S, int HW) {
    CNB_PDL_SYNC();
    const long total = (long)B * C;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        const int c = (int)(i % C), b = (int)(i / C);
        float sum = 0.f, mx = -INFINITY;
        int arg = 0;
        for (int s = 0; s < S; ++s) {
            const long o = ((long)b * S + s) * C + c;
            sum += part_sum[o];
            if (part_max[o] > mx) mx = part_max[o], arg = part_arg[o];  // slices ascend: a strict test keeps the first maximum
        }
        ch_avg[i] = sum / (float)HW;
        ch_max[i] = mx;
        ch_arg[i] = arg;
    }
}

// dx[p][c] = dsp[p].mean / C + [x == max_p] dsp[p].max / ties_p + dch_avg[b][c] / HW + [p == arg_bc] dch_max[b][c]
template <typename T, int VEC>
__global__ void __launch_bounds__(256) sca_pool_bwd_kernel(const T* __restrict__ x, const float* __restrict__ sp,
                                                          const float* __restrict__ ties, const float* __restrict__ dsp,
                                                          const float* __restrict__ dch_avg, const float* __restrict__ dch_max,
                                                          const int* __restrict__ ch_arg, T* __restrict__ dx, int B, int HW, int C) {
    CNB_PDL_SYNC();
    const int CV = C / VEC;
    const long total = (long)B * HW * CV;
    const float inv_c = 1.0f / (float)C, inv_hw = 1.0f / (float)HW;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        const int g = (int)(i % CV);
        const long p = i / CV;
        const int pl = (int)(p % HW), b = (int)(p / HW);
        float v[VEC], r[VEC];
        pa_ld<T, VEC>(x + p * C + g * VEC, v);
        const float dmean = dsp[2 * p] * inv_c, mx = sp[2 * p + 1], dmax = dsp[2 * p + 1] / ties[p];
#pragma unroll
        for (int j = 0; j < VEC; ++j) {
            const int c = g * VEC + j;
            float t = dmean + (v[j] == mx ? dmax : 0.f) + dch_avg[(long)b * C + c] * inv_hw;
            if (ch_arg[(long)b * C + c] == pl) t += dch_max[(long)b * C + c];
            r[j] = t;
        }
        pa_st<T, VEC>(dx + p * C + g * VEC, r);
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// spatial-channel attention, apply side (attention.py:118-123 + convolution.py:392-393):
//   out = y * (1 + gamma * 0.5 * (sigmoid(cl[b][c]) + sigmoid(sl[p])))
// ---------------------------------------------------------------------------------------------------------------------
template <typename T, int VEC>
__global__ void __launch_bounds__(256) sca_apply_fwd_kernel(const T* __restrict__ y, const float* __restrict__ cl, const float* __restrict__ sl,
                                                           const float* __restrict__ gamma, T* __restrict__ out, int B, int HW, int C) {
    CNB_PDL_SYNC();
    const int CV = C / VEC;
    const long total = (long)B * HW * CV;
    const float hg = 0.5f * gamma[0];
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        const int g = (int)(i % CV);
        const long p = i / CV;
        const int b = (int)(p / HW);
        const float ss = cnb_sigmoid(sl[p]);
        float v[VEC];
        pa_ld<T, VEC>(y + p * C + g * VEC, v);
#pragma unroll
        for (int j = 0; j < VEC; ++j) v[j] *= 1.0f + hg * (cnb_sigmoid(cl[(long)b * C + g * VEC + j]) + ss);
        pa_st<T, VEC>(out + p * C + g * VEC, v);
    }
}

// grid (S slices, B samples, column tiles), same thread map as the pooling kernel.  dy is written; dcl[b][c], dsl[p] and dgamma are
// accumulated with fp32 atomics into zeroed buffers (dsl after a segmented warp reduction when the column count allows it).
template <typename T, int VEC>
__global__ void __launch_bounds__(SCA_THREADS) sca_apply_bwd_kernel(const T* __restrict__ y, const T* __restrict__ dout,
                                                                   const float* __restrict__ cl, const float* __restrict__ sl,
                                                                   const float* __restrict__ gamma, T* __restrict__ dy, float* __restrict__ dcl,
                                                                   float* __restrict__ dsl, float* __restrict__ dgamma, int HW, int C, int S) {
    CNB_PDL_SYNC();
    __shared__ float sh_red[SCA_THREADS / 32];
    const int CV = C / VEC;
    const int cols = CV < SCA_THREADS ? CV : SCA_THREADS;
    const int R = SCA_THREADS / cols;
    const int r = threadIdx.x / cols, gl = threadIdx.x % cols;
    const int g = blockIdx.z * cols + gl;
    const bool active = r < R && g < CV;
    const int s = blockIdx.x, b = blockIdx.y;
    const int chunk = (HW + S - 1) / S;
    const int p0 = s * chunk, p1 = p0 + chunk < HW ? p0 + chunk : HW;
    // lanes that share a pixel: a power-of-two segment of the warp (cols >= 32 and a multiple of 32: the whole warp)
    const bool pow2 = (cols & (cols - 1)) == 0;
    const int seg = (cols % 32 == 0) ? 32 : ((pow2 && cols < 32) ? cols : 1);
    const float hg = 0.5f * gamma[0];
    float sc[VEC], dcl_acc[VEC];
#pragma unroll
    for (int j = 0; j < VEC; ++j) {
        sc[j] = active ? cnb_sigmoid(cl[(long)b * C + g * VEC + j]) : 0.f;
        dcl_acc[j] = 0.f;
    }
    float dg_acc = 0.f;
    for (int pb = p0; pb < p1; pb += R) {  // uniform trip count: every lane takes part in the shuffles
        const int p = pb + r;
        const bool ok = active && p < p1;
        float part = 0.f;
        if (ok) {
            const long row = (long)b * HW + p;
            const float ss = cnb_sigmoid(sl[row]);
            float yv[VEC], gv[VEC], o[VEC];
            pa_ld<T, VEC>(y + row * C + g * VEC, yv);
            pa_ld<T, VEC>(dout + row * C + g * VEC, gv);
            float tsum = 0.f;
#pragma unroll
            for (int j = 0; j < VEC; ++j) {
                const float t = gv[j] * yv[j];  // d(out)/d(att) contribution
                o[j] = gv[j] * (1.0f + hg * (sc[j] + ss));
                dcl_acc[j] += t;
                dg_acc += t * (sc[j] + ss);
                tsum += t;
            }
            pa_st<T, VEC>(dy + row * C + g * VEC, o);
            part = tsum * hg * ss * (1.0f - ss);
        }
        for (int o2 = seg / 2; o2 > 0; o2 >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o2);
        if (ok && (gl % seg) == 0) atomicAdd(dsl + (long)b * HW + p, part);
    }
    if (active) {
#pragma unroll
        for (int j = 0; j < VEC; ++j) atomicAdd(dcl + (long)b * C + g * VEC + j, dcl_acc[j] * hg * sc[j] * (1.0f - sc[j]));
    }
    dg_acc = cnb_warp_sum(dg_acc);
    if ((threadIdx.x & 31) == 0) sh_red[threadIdx.x >> 5] = dg_acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.f;
        for (int w = 0; w < SCA_THREADS / 32; ++w) t += sh_red[w];
        atomicAdd(dgamma, 0.5f * t);
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// dropout with a counter-based generator.  state = (seed, step counter) in device memory, so a captured CUDA graph draws a new mask
// at every replay (cnb_rng_advance bumps the counter once per forward); `site` separates the call sites of one step.  One 64-bit
// hash serves four elements (16 bits each): keep iff bits >= thr = round(p * 65536).  The backward re-derives the same mask.
// ---------------------------------------------------------------------------------------------------------------------
__global__ void rng_advance_kernel(int64_t* state) {
    CNB_PDL_SYNC();
    if (blockIdx.x == 0 && threadIdx.x == 0) state[1] += 1;
}

// elementwise (nn.Dropout): out = x * keep / (1 - p)
template <typename T, int VEC>
__global__ void __launch_bounds__(256) dropout_kernel(const T* __restrict__ x, T* __restrict__ out, long n_v, const int64_t* __restrict__ state,
                                                     int site, uint32_t thr, float scale) {
    CNB_PDL_SYNC();
    const uint64_t key = cnb_rng_key(state, site);
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n_v; i += (long)gridDim.x * blockDim.x) {
        float v[VEC];
        pa_ld<T, VEC>(x + i * VEC, v);
#pragma unroll
        for (int j = 0; j < VEC; ++j) v[j] = cnb_rng_bits16(key, (uint64_t)(i * VEC + j)) >= thr ? v[j] * scale : 0.f;
        pa_st<T, VEC>(out + i * VEC, v);
    }
}

// channel-wise (nn.Dropout2d): one draw per (sample, channel)
template <typename T, int VEC>
__global__ void __launch_bounds__(256) dropout2d_kernel(const T* __restrict__ x, T* __restrict__ out, int B, int HW, int C,
                                                       const int64_t* __restrict__ state, int site, uint32_t thr, float scale) {
    CNB_PDL_SYNC();
    const uint64_t key = cnb_rng_key(state, site);
    const int CV = C / VEC;
    const long total = (long)B * HW * CV;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        const int g = (int)(i % CV);
        const long p = i / CV;
        const int b = (int)(p / HW);
        float v[VEC];
        pa_ld<T, VEC>(x + p * C + g * VEC, v);
#pragma unroll
        for (int j = 0; j < VEC; ++j) v[j] = cnb_rng_bits16(key, (uint64_t)((long)b * C + g * VEC + j)) >= thr ? v[j] * scale : 0.f;
        pa_st<T, VEC>(out + p * C + g * VEC, v);
    }
}

}  // namespace cnb
