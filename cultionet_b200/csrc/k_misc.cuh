// Bandwidth-bound pieces of the TowerUNet path: bilinear fix-up resize, PreTimeReduction's temporal
// convolution, and the fused TowerUNetFinalCombine / SigmoidCrisp head.
#pragma once
#include "cnb_common.cuh"

namespace cnb {

// ---------------------------------------------------------------------------------------------
// bilinear, align_corners=True (ATen upsample_bilinear2d semantics: src = dst * (in-1)/(out-1))
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void bilinear_src(int o, float rscale, int in_len, int& i0, int& i1, float& l0, float& l1) {
    const float r = rscale * (float)o;
    i0 = (int)r;
    if (i0 > in_len - 1) i0 = in_len - 1;
    i1 = i0 + ((i0 < in_len - 1) ? 1 : 0);
    l1 = r - (float)i0;
    l0 = 1.f - l1;
}

template <typename T>
__global__ void __launch_bounds__(256) resize_bilinear_fwd_kernel(const T* __restrict__ x, T* __restrict__ y, int B, int Hin, int Win,
                                                                 int Hout, int Wout, int C, float rh, float rw) {
    CNB_PDL_SYNC();
    const long total = (long)B * Hout * Wout * C;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        const int c = (int)(i % C);
        long t = i / C;
        const int ox = (int)(t % Wout);
        t /= Wout;
        const int oy = (int)(t % Hout);
        const long b = t / Hout;
        int y0, y1, x0, x1;
        float ly0, ly1, lx0, lx1;
        bilinear_src(oy, rh, Hin, y0, y1, ly0, ly1);
        bilinear_src(ox, rw, Win, x0, x1, lx0, lx1);
        const T* base = x + b * Hin * Win * C + c;
        const float v00 = cnb_ld(base + ((long)y0 * Win + x0) * C);
        const float v01 = cnb_ld(base + ((long)y0 * Win + x1) * C);
        const float v10 = cnb_ld(base + ((long)y1 * Win + x0) * C);
        const float v11 = cnb_ld(base + ((long)y1 * Win + x1) * C);
        cnb_st(y + i, ly0 * (lx0 * v00 + lx1 * v01) + ly1 * (lx0 * v10 + lx1 * v11));
    }
}

// weight with which output index `o` reads input index `i` along one axis
__device__ __forceinline__ float bilinear_weight(int o, int i, float rscale, int in_len) {
    int i0, i1;
    float l0, l1;
    bilinear_src(o, rscale, in_len, i0, i1, l0, l1);
    float w = 0.f;
    if (i0 == i) w += l0;
    if (i1 == i) w += l1;
    return w;
}

__device__ __forceinline__ void bilinear_candidates(int i, float rscale, int out_len, int& lo, int& hi) {
    if (rscale <= 0.f) {
        lo = 0;
        hi = out_len - 1;
        return;
    }
    lo = (int)floorf(((float)i - 1.f) / rscale) - 1;
    hi = (int)ceilf(((float)i + 1.f) / rscale) + 1;
    if (lo < 0) lo = 0;
    if (hi > out_len - 1) hi = out_len - 1;
}

// gather form of the adjoint: deterministic, no atomics
template <typename T>
__global__ void __launch_bounds__(256) resize_bilinear_bwd_kernel(const T* __restrict__ dy, T* __restrict__ dx, int B, int Hin, int Win,
                                                                 int Hout, int Wout, int C, float rh, float rw) {
    CNB_PDL_SYNC();
    const long total = (long)B * Hin * Win * C;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        const int c = (int)(i % C);
        long t = i / C;
        const int ix = (int)(t % Win);
        t /= Win;
        const int iy = (int)(t % Hin);
        const long b = t / Hin;
        int ylo, yhi, xlo, xhi;
        bilinear_candidates(iy, rh, Hout, ylo, yhi);
        bilinear_candidates(ix, rw, Wout, xlo, xhi);
        const T* base = dy + b * Hout * Wout * C + c;
        float acc = 0.f;
        for (int oy = ylo; oy <= yhi; ++oy) {
            const float wy = bilinear_weight(oy, iy, rh, Hin);
            if (wy == 0.f) continue;
            float row = 0.f;
            for (int ox = xlo; ox <= xhi; ++ox) {
                const float wx = bilinear_weight(ox, ix, rw, Win);
                if (wx != 0.f) row = fmaf(wx, cnb_ld(base + ((long)oy * Wout + ox) * C), row);
            }
            acc = fmaf(wy, row, acc);
        }
        cnb_st(dx + i, acc);
    }
}

// ---------------------------------------------------------------------------------------------
// PreTimeReduction stage 1: valid temporal convolution C -> C with kernel k over x[B,C,T,H,W] (fp32).
// Output u[p][c2*T' + t'] pixel-major with a pixel pitch `up` >= C*T' (padding columns are written as zero), so that stage 2 is a
// plain 1x1 GEMM with K = C*T' that the TMA-fed kernels can read.
//
// One CTA = PT_PIX consecutive pixels.  x is channel/time-major, so for a fixed (c, t) those pixels are contiguous floats:
// the tile is staged in shared memory with coalesced loads, every thread then owns one pixel and a quarter of the outputs
// (bank = pixel: conflict-free; the filter taps are warp-uniform broadcasts), the results go back through shared memory and
// leave as whole 16-byte vectors of the [pixels][pitch] output block.  (The first version wrote 2-byte elements 220 bytes
// apart and ran at 1 TB/s, ncu r01i.)
// ---------------------------------------------------------------------------------------------
constexpr int PT_PIX = 64;
constexpr int PT_THREADS = 256;
constexpr int PT_MAX_K = 8;

// dynamic smem: xs[C*T][PT_PIX] floats, then the output tile us[PT_PIX][up] of T, then w[C*C*k] floats
template <typename T>
__global__ void __launch_bounds__(PT_THREADS) pretime_conv_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w1,
                                                                     T* __restrict__ u, int B, int C, int Tn, int H, int W, int k, int up) {
    CNB_PDL_SYNC();
    CNB_DYN_SMEM(sm_raw);
    const int Tp = Tn - k + 1;
    const int CT = C * Tn;
    float* xs = reinterpret_cast<float*>(sm_raw);
    T* us = reinterpret_cast<T*>(xs + (long)CT * PT_PIX);
    float* ws = reinterpret_cast<float*>(us + (long)PT_PIX * up);
    const long HW = (long)H * W;
    const long total = (long)B * HW;
    for (int i = threadIdx.x; i < C * C * k; i += blockDim.x) ws[i] = w1[i];
    const int p_in = threadIdx.x % PT_PIX, grp = threadIdx.x / PT_PIX, ngrp = blockDim.x / PT_PIX;
    for (long p0 = (long)blockIdx.x * PT_PIX; p0 < total; p0 += (long)gridDim.x * PT_PIX) {
        __syncthreads();  // previous tile fully written out (and ws visible on the first pass)
        // ---- stage x: row (c, t), PT_PIX pixels
        for (int i = threadIdx.x; i < CT * PT_PIX; i += blockDim.x) {
            const int pp = i % PT_PIX, ct = i / PT_PIX;
            const long p = p0 + pp;
            float v = 0.f;
            if (p < total) {
                const long b = p / HW, hw = p - b * HW;
                v = x[(b * CT + ct) * HW + hw];
            }
            xs[ct * PT_PIX + pp] = v;
        }
        __syncthreads();
        // ---- compute: thread = (pixel, output group)
        for (int j = grp; j < up; j += ngrp) {
            float acc = 0.f;
            if (j < C * Tp) {
                const int c2 = j / Tp, tp = j - c2 * Tp;
                for (int c = 0; c < C; ++c) {
                    const float* xr = xs + (c * Tn + tp) * PT_PIX + p_in;
                    const float* wr = ws + (c2 * C + c) * k;
                    for (int dt = 0; dt < k; ++dt) acc = fmaf(wr[dt], xr[dt * PT_PIX], acc);
                }
            }
            cnb_st(us + (long)p_in * up + j, acc);
        }
        __syncthreads();
        // ---- write the [PT_PIX][up] block: contiguous in global memory
        const long valid_px = total - p0 < PT_PIX ? total - p0 : PT_PIX;
        const long nelem = valid_px * up;
        T* dst = u + p0 * up;
        constexpr int V = cnb_vec<T>::N;
        if (up % V == 0 && cnb_aligned16_dev(dst)) {
            for (long i = threadIdx.x; i < nelem / V; i += blockDim.x)
                *reinterpret_cast<uint4*>(dst + i * V) = *reinterpret_cast<const uint4*>(us + i * V);
        } else {
            for (long i = threadIdx.x; i < nelem; i += blockDim.x) dst[i] = us[i];
        }
    }
}

// dw1[c2][c][dt] += sum_{p,t'} du[p][c2*T'+t'] * x[b,c,t'+dt,hw].  Same tiling; every thread owns one pixel of the tile and a
// quarter of the (c2, c, dt) triples, keeps its partial sums in registers across all tiles of the CTA, and the CTA reduces once.
constexpr int PT_WG_MAX_TRIPLES = 64;  // per thread: ceil(C*C*k / 4) must fit
template <typename T>
__global__ void __launch_bounds__(PT_THREADS) pretime_conv_wgrad_kernel(const float* __restrict__ x, const T* __restrict__ du,
                                                                       float* __restrict__ dw1, int B, int C, int Tn, int H, int W, int k,
                                                                       int up) {
    CNB_PDL_SYNC();
    CNB_DYN_SMEM(sm_raw);
    const int Tp = Tn - k + 1;
    const int CT = C * Tn;
    float* xs = reinterpret_cast<float*>(sm_raw);
    T* ds = reinterpret_cast<T*>(xs + (long)CT * PT_PIX);   // [PT_PIX][up + 2]: odd word pitch against bank conflicts
    const int dpitch = up + 2;
    float* red = reinterpret_cast<float*>(ds + (long)PT_PIX * dpitch);  // [C*C*k]
    const long HW = (long)H * W;
    const long total = (long)B * HW;
    const int ntr = C * C * k;
    const int p_in = threadIdx.x % PT_PIX, grp = threadIdx.x / PT_PIX, ngrp = blockDim.x / PT_PIX;
    float acc[PT_WG_MAX_TRIPLES];
#pragma unroll
    for (int i = 0; i < PT_WG_MAX_TRIPLES; ++i) acc[i] = 0.f;
    for (int i = threadIdx.x; i < ntr; i += blockDim.x) red[i] = 0.f;
    for (long p0 = (long)blockIdx.x * PT_PIX; p0 < total; p0 += (long)gridDim.x * PT_PIX) {
        __syncthreads();
        for (int i = threadIdx.x; i < CT * PT_PIX; i += blockDim.x) {
            const int pp = i % PT_PIX, ct = i / PT_PIX;
            const long p = p0 + pp;
            float v = 0.f;
            if (p < total) {
                const long b = p / HW, hw = p - b * HW;
                v = x[(b * CT + ct) * HW + hw];
            }
            xs[ct * PT_PIX + pp] = v;
        }
        for (int i = threadIdx.x; i < PT_PIX * up; i += blockDim.x) {
            const int pp = i / up, j = i - pp * up;
            ds[pp * dpitch + j] = (p0 + pp < total) ? du[(p0 + pp) * up + j] : T(0.f);
        }
        __syncthreads();
#pragma unroll
        for (int i = 0; i < PT_WG_MAX_TRIPLES; ++i) {
            const int tr = grp + i * ngrp;
            if (tr < ntr) {
                const int dt = tr % k, cc = tr / k;
                const int c = cc % C, c2 = cc / C;
                const float* xr = xs + (c * Tn + dt) * PT_PIX + p_in;
                const T* dr = ds + p_in * dpitch + c2 * Tp;
                float a = acc[i];
                for (int tp = 0; tp < Tp; ++tp) a = fmaf(cnb_ld(dr + tp), xr[tp * PT_PIX], a);
                acc[i] = a;
            }
        }
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < PT_WG_MAX_TRIPLES; ++i) {
        const int tr = grp + i * ngrp;
        if (tr < ntr) {  // uniform per warp: a warp holds 32 pixels of one group
            const float v = cnb_warp_sum(acc[i]);
            if ((threadIdx.x & 31) == 0) atomicAdd(&red[tr], v);
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < ntr; i += blockDim.x) atomicAdd(dw1 + i, red[i]);
}

// ---------------------------------------------------------------------------------------------
// PreTimeReduction stage 1 as a GEMM (throughput mode).  The valid temporal convolution is a banded (Toeplitz) matrix product
//   u[p][c2*T' + t'] = sum_{c,t} xp[p][c*T + t] * Wt[c2*T' + t'][c*T + t],   Wt[..][..] = w1[c2][c][t - t'] for 0 <= t - t' < k
// so once x is pixel-major it is a 1x1 convolution with K = C*T that the tcgen05 kernels run in a few tens of microseconds
// (13.8 GFLOP dense at cfg 2 instead of a shared-memory-bound SIMT loop), and ONE transposed copy of x serves both temporal
// branches.  Three small kernels: the transpose, the expansion of w1 into Wt and the fold of dWt back onto dw1.
// ---------------------------------------------------------------------------------------------
constexpr int TP_PIX = 64;
constexpr int TP_XPITCH = TP_PIX + 1;

// x[B][CT][HW] fp32 -> xp[B*HW][pitch] (columns >= CT are zero).  dynamic smem: xs[CT][TP_XPITCH] floats, us[TP_PIX][pitch] of T
template <typename T>
__global__ void __launch_bounds__(256) time_to_pixel_major_kernel(const float* __restrict__ x, T* __restrict__ xp, int B, int CT, long HW,
                                                                 int pitch) {
    CNB_PDL_SYNC();
    constexpr int V = cnb_vec<T>::N;
    CNB_DYN_SMEM(sm_raw);
    float* xs = reinterpret_cast<float*>(sm_raw);
    T* us = reinterpret_cast<T*>(xs + ((long)CT * TP_XPITCH + 3) / 4 * 4);  // 16-byte aligned: whole-vector copies below
    const long total = (long)B * HW;
    const int groups = pitch / V;
    for (long p0 = (long)blockIdx.x * TP_PIX; p0 < total; p0 += (long)gridDim.x * TP_PIX) {
        __syncthreads();  // the previous tile has left shared memory
        if (HW % TP_PIX == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0) {
            // the tile lies inside one image: one division per tile, 16-byte loads (four pixels of one (channel, time) plane)
            const long b = p0 / HW, hw0 = p0 - b * HW;
            const float* xb = x + b * CT * HW + hw0;
            for (int i = threadIdx.x; i < CT * (TP_PIX / 4); i += blockDim.x) {
                const int q = i % (TP_PIX / 4), ct = i / (TP_PIX / 4);
                const float4 v = *reinterpret_cast<const float4*>(xb + (long)ct * HW + q * 4);
                float* d = xs + ct * TP_XPITCH + q * 4;
                d[0] = v.x, d[1] = v.y, d[2] = v.z, d[3] = v.w;
            }
        } else {
            for (int i = threadIdx.x; i < CT * TP_PIX; i += blockDim.x) {
                const int pp = i % TP_PIX, ct = i / TP_PIX;
                const long p = p0 + pp;
                float v = 0.f;
                if (p < total) {
                    const long b = p / HW, hw = p - b * HW;
                    v = x[(b * CT + ct) * HW + hw];
                }
                xs[ct * TP_XPITCH + pp] = v;
            }
        }
        __syncthreads();
        // thread = (pixel, column group): lanes walk pixels, so the xs reads are conflict-free
        for (int i = threadIdx.x; i < groups * TP_PIX; i += blockDim.x) {
            const int pp = i % TP_PIX, g = i / TP_PIX;
            float v[V];
#pragma unroll
            for (int j = 0; j < V; ++j) {
                const int ct = g * V + j;
                v[j] = ct < CT ? xs[ct * TP_XPITCH + pp] : 0.f;
            }
            cnb_stv(us + (long)pp * pitch + g * V, v);
        }
        __syncthreads();
        const long valid_px = total - p0 < TP_PIX ? total - p0 : TP_PIX;
        const long nvec = valid_px * groups;
        T* dst = xp + p0 * pitch;
        for (long i = threadIdx.x; i < nvec; i += blockDim.x)
            *reinterpret_cast<uint4*>(dst + i * V) = *reinterpret_cast<const uint4*>(us + i * V);
    }
}

// Wt[n][c*T + t] (n < rows; rows past C*T' are zero)
__global__ void __launch_bounds__(256) toeplitz_expand_kernel(const float* __restrict__ w1, float* __restrict__ wt, int C, int Tn, int k,
                                                             int rows) {
    CNB_PDL_SYNC();
    const int Tp = Tn - k + 1, CT = C * Tn;
    const int total = rows * CT;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int n = i / CT, kk = i - n * CT;
        const int c2 = n / Tp, tp = n - c2 * Tp;
        const int c = kk / Tn, t = kk - c * Tn;
        const int dt = t - tp;
        wt[i] = (c2 < C && dt >= 0 && dt < k) ? w1[(c2 * C + c) * k + dt] : 0.f;
    }
}

// dw1[c2][c][dt] = sum_t' dWt[c2*T' + t'][c*T + t' + dt]
__global__ void __launch_bounds__(256) toeplitz_fold_kernel(const float* __restrict__ dwt, float* __restrict__ dw1, int C, int Tn, int k) {
    CNB_PDL_SYNC();
    const int Tp = Tn - k + 1, CT = C * Tn;
    const int total = C * C * k;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int dt = i % k, cc = i / k;
        const int c = cc % C, c2 = cc / C;
        float acc = 0.f;
        for (int tp = 0; tp < Tp; ++tp) acc += dwt[(long)(c2 * Tp + tp) * CT + c * Tn + tp + dt];
        dw1[i] = acc;
    }
}

// ---------------------------------------------------------------------------------------------
// Skinny 3x3 convolutions (the Psi-Net stream heads, 256 -> 9): out[p][n] = sum_tap sum_c x[p + off(tap)][c] w[n][c][tap].
// As an implicit GEMM with N = 9 the tensor-core kernel re-reads the 256-channel input once per tap from L2 (2.4 GB per call, 0.29 ms,
// L2-bound).  Instead: t[p][(tap, n)] = sum_c x[p][c] w[n][c][tap] is ONE 1x1 GEMM with N = taps*n (81 -> 88) that reads x once, and
// the kernels below do the remaining shift-and-add over the narrow t tensor / its adjoint gather (which is also exactly the im2col
// of dy that the data- and weight-gradient GEMMs of the 1x1 form need).  unit stride, zero padding.
// ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) tap_shift_add_kernel(const T* __restrict__ t, T* __restrict__ out, int B, int H, int W, int N, int KH,
                                                           int KW, int pad, int dil, int t_pitch, int out_pitch) {
    CNB_PDL_SYNC();
    const long P = (long)B * H * W;
    for (long p = (long)blockIdx.x * blockDim.x + threadIdx.x; p < P; p += (long)gridDim.x * blockDim.x) {
        const int x = (int)(p % W);
        const int y = (int)((p / W) % H);
        float acc[16];
#pragma unroll
        for (int n = 0; n < 16; ++n) acc[n] = 0.f;
        for (int ky = 0; ky < KH; ++ky) {
            const int iy = y + ky * dil - pad;
            if (iy < 0 || iy >= H) continue;
            for (int kx = 0; kx < KW; ++kx) {
                const int ix = x + kx * dil - pad;
                if (ix < 0 || ix >= W) continue;
                const T* src = t + (p + (long)(iy - y) * W + (ix - x)) * t_pitch + (ky * KW + kx) * N;
#pragma unroll
                for (int n = 0; n < 16; ++n)
                    if (n < N) acc[n] += cnb_ld(src + n);
            }
        }
#pragma unroll
        for (int n = 0; n < 16; ++n)
            if (n < N) cnb_st(out + p * out_pitch + n, acc[n]);
    }
}

// adjoint: dt[q][(tap, n)] = dout[q - off(tap)][n] (zero outside the image; columns >= taps*N zero)
template <typename T>
__global__ void __launch_bounds__(256) tap_shift_gather_kernel(const T* __restrict__ dout, T* __restrict__ dt, int B, int H, int W, int N,
                                                              int KH, int KW, int pad, int dil, int t_pitch, int out_pitch) {
    CNB_PDL_SYNC();
    const long total = (long)B * H * W * t_pitch;
    const int taps = KH * KW;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        const int col = (int)(i % t_pitch);
        const long q = i / t_pitch;
        float v = 0.f;
        if (col < taps * N) {
            const int tap = col / N, n = col - tap * N;
            const int ky = tap / KW, kx = tap - ky * KW;
            const int x = (int)(q % W);
            const int y = (int)((q / W) % H);
            const int oy = y - (ky * dil - pad), ox = x - (kx * dil - pad);
            if (oy >= 0 && oy < H && ox >= 0 && ox < W) v = cnb_ld(dout + (q + (long)(oy - y) * W + (ox - x)) * out_pitch + n);
        }
        cnb_st(dt + i, v);
    }
}

// The same gather with one thread per PIXEL for a compile-time (kernel size, N): the generic kernel above spends ~6 runtime integer
// divisions per ELEMENT (305 us for the [32,128,128,88] bf16 tensor of the Psi-Net heads, 14 us of HBM time); here a thread assembles
// its whole row of KS*KS*NN columns in registers (KS*KS*NN short loads that hit L1) and writes it with 16-byte stores.
template <typename T, int NN, int KS>
__global__ void __launch_bounds__(256) tap_shift_gather_px_kernel(const T* __restrict__ dout, T* __restrict__ dt, int B, int H, int W, int pad,
                                                                 int dil, int t_pitch, int out_pitch) {
    CNB_PDL_SYNC();
    constexpr int V = cnb_vec<T>::N;
    constexpr int COLS = KS * KS * NN;
    constexpr int ROW = (COLS + V - 1) / V * V;
    const long P = (long)B * H * W;
    for (long p = (long)blockIdx.x * blockDim.x + threadIdx.x; p < P; p += (long)gridDim.x * blockDim.x) {
        const int x = (int)(p % W);
        const int y = (int)((p / W) % H);
        float row[ROW];
#pragma unroll
        for (int ky = 0; ky < KS; ++ky)
#pragma unroll
            for (int kx = 0; kx < KS; ++kx) {
                const int oy = y - (ky * dil - pad), ox = x - (kx * dil - pad);
                const bool ok = oy >= 0 && oy < H && ox >= 0 && ox < W;
                const T* src = dout + (p + (long)(oy - y) * W + (ox - x)) * out_pitch;
#pragma unroll
                for (int n = 0; n < NN; ++n) row[(ky * KS + kx) * NN + n] = ok ? cnb_ld(src + n) : 0.f;
            }
#pragma unroll
        for (int j = COLS; j < ROW; ++j) row[j] = 0.f;
        T* dst = dt + p * t_pitch;
#pragma unroll
        for (int j = 0; j < ROW; j += V) cnb_stv(dst + j, row + j);
        for (int j = ROW; j < t_pitch; ++j) cnb_st(dst + j, 0.f);
    }
}

// ---------------------------------------------------------------------------------------------
// TowerUNetFinalCombine (+ SigmoidCrisp).  params: g[3][3], w[3], b[3], crisp_gamma
// ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) final_combine_fwd_kernel(const T* __restrict__ ha, const T* __restrict__ hb,
                                                               const T* __restrict__ hc, const float* __restrict__ prm, float smooth,
                                                               int flags, float* __restrict__ dist, float* __restrict__ edge,
                                                               float* __restrict__ crop, long P) {
    CNB_PDL_SYNC();
    float ig[9], w[3], bb[3];
#pragma unroll
    for (int i = 0; i < 9; ++i) ig[i] = 1.0f / prm[i];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        w[i] = prm[9 + i];
        bb[i] = prm[12 + i];
    }
    const float r = 1.0f / (smooth + cnb_sigmoid(prm[15]));
    for (long p = (long)blockIdx.x * blockDim.x + threadIdx.x; p < P; p += (long)gridDim.x * blockDim.x) {
        float z[3];
#pragma unroll
        for (int t = 0; t < 3; ++t) {
            const float s = cnb_ld(ha + p * 3 + t) * ig[t * 3 + 0] + cnb_ld(hb + p * 3 + t) * ig[t * 3 + 1] +
                            cnb_ld(hc + p * 3 + t) * ig[t * 3 + 2];
            z[t] = fmaf(w[t], s, bb[t]);
        }
        dist[p] = cnb_sigmoid(z[0]);
        edge[p] = (flags & 1) ? cnb_sigmoid(z[1] * r) : z[1];
        crop[p] = (flags & 2) ? cnb_sigmoid(z[2]) : z[2];
    }
}

// red[0..8] = A_tj = sum dz_t*h_j ; red[9..11] = sum dz_t ; red[12] = sum du*z_1 (SigmoidCrisp scale)
template <typename T>
__global__ void __launch_bounds__(256) final_combine_bwd_kernel(const T* __restrict__ ha, const T* __restrict__ hb,
                                                               const T* __restrict__ hc, const float* __restrict__ prm, float smooth,
                                                               int flags, const float* __restrict__ d_dist,
                                                               const float* __restrict__ d_edge, const float* __restrict__ d_crop,
                                                               T* __restrict__ dha, T* __restrict__ dhb, T* __restrict__ dhc,
                                                               float* __restrict__ red, long P) {
    CNB_PDL_SYNC();
    __shared__ float sh[13];
    if (threadIdx.x < 13) sh[threadIdx.x] = 0.f;
    __syncthreads();
    float ig[9], w[3], bb[3];
#pragma unroll
    for (int i = 0; i < 9; ++i) ig[i] = 1.0f / prm[i];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        w[i] = prm[9 + i];
        bb[i] = prm[12 + i];
    }
    const float r = 1.0f / (smooth + cnb_sigmoid(prm[15]));
    float acc[13];
#pragma unroll
    for (int i = 0; i < 13; ++i) acc[i] = 0.f;
    for (long p = (long)blockIdx.x * blockDim.x + threadIdx.x; p < P; p += (long)gridDim.x * blockDim.x) {
        float h[3][3], z[3], dz[3];
#pragma unroll
        for (int t = 0; t < 3; ++t) {
            h[t][0] = cnb_ld(ha + p * 3 + t);
            h[t][1] = cnb_ld(hb + p * 3 + t);
            h[t][2] = cnb_ld(hc + p * 3 + t);
            z[t] = fmaf(w[t], h[t][0] * ig[t * 3] + h[t][1] * ig[t * 3 + 1] + h[t][2] * ig[t * 3 + 2], bb[t]);
        }
        {
            const float o = cnb_sigmoid(z[0]);
            dz[0] = d_dist[p] * o * (1.f - o);
        }
        if (flags & 1) {
            const float o = cnb_sigmoid(z[1] * r);
            const float du = d_edge[p] * o * (1.f - o);
            dz[1] = du * r;
            acc[12] = fmaf(du, z[1], acc[12]);
        } else {
            dz[1] = d_edge[p];
        }
        if (flags & 2) {
            const float o = cnb_sigmoid(z[2]);
            dz[2] = d_crop[p] * o * (1.f - o);
        } else {
            dz[2] = d_crop[p];
        }
#pragma unroll
        for (int t = 0; t < 3; ++t) {
            cnb_st(dha + p * 3 + t, dz[t] * w[t] * ig[t * 3 + 0]);
            cnb_st(dhb + p * 3 + t, dz[t] * w[t] * ig[t * 3 + 1]);
            cnb_st(dhc + p * 3 + t, dz[t] * w[t] * ig[t * 3 + 2]);
#pragma unroll
            for (int j = 0; j < 3; ++j) acc[t * 3 + j] = fmaf(dz[t], h[t][j], acc[t * 3 + j]);
            acc[9 + t] += dz[t];
        }
    }
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int i = 0; i < 13; ++i) {
        const float s = cnb_warp_sum(acc[i]);
        if (lane == 0) atomicAdd(&sh[i], s);
    }
    __syncthreads();
    if (threadIdx.x < 13) atomicAdd(red + threadIdx.x, sh[threadIdx.x]);
}

__global__ void final_combine_param_grad_kernel(const float* __restrict__ prm, const float* __restrict__ red, float smooth, int flags,
                                                float* __restrict__ dprm) {
    CNB_PDL_SYNC();
    const int i = threadIdx.x;
    if (i < 9) {
        const int t = i / 3;
        const float g = prm[i];
        dprm[i] = -prm[9 + t] * red[i] / (g * g);
    } else if (i < 12) {
        const int t = i - 9;
        dprm[i] = red[t * 3] / prm[t * 3] + red[t * 3 + 1] / prm[t * 3 + 1] + red[t * 3 + 2] / prm[t * 3 + 2];
    } else if (i < 15) {
        dprm[i] = red[9 + (i - 12)];
    } else if (i == 15) {
        if (flags & 1) {
            const float sg = cnb_sigmoid(prm[15]);
            const float r = 1.0f / (smooth + sg);
            dprm[15] = red[12] * (-r * r) * sg * (1.f - sg);
        } else {
            dprm[15] = 0.f;
        }
    }
}

}  // namespace cnb
