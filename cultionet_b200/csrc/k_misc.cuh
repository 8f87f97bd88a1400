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
// PreTimeReduction stage 1: valid temporal convolution C -> C with kernel k over x[B,C,T,H,W] (fp32)
// output u[p][c2*T' + t'] pixel-major so that stage 2 is a plain 1x1 GEMM with K = C*T'
// ---------------------------------------------------------------------------------------------
constexpr int PT_CHUNK = 8;

template <typename T>
__global__ void __launch_bounds__(256) pretime_conv_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w1,
                                                              T* __restrict__ u, int B, int C, int Tn, int H, int W, int k) {
    const int Tp = Tn - k + 1;
    const long HW = (long)H * W;
    const long total = (long)B * HW;
    const int c2 = blockIdx.y;
    for (long p = (long)blockIdx.x * blockDim.x + threadIdx.x; p < total; p += (long)gridDim.x * blockDim.x) {
        const long b = p / HW, hw = p - b * HW;
        const float* xb = x + b * C * Tn * HW + hw;
        T* up = u + p * ((long)C * Tp) + (long)c2 * Tp;
        for (int t0 = 0; t0 < Tp; t0 += PT_CHUNK) {
            float acc[PT_CHUNK];
#pragma unroll
            for (int j = 0; j < PT_CHUNK; ++j) acc[j] = 0.f;
            for (int c = 0; c < C; ++c) {
                const float* xc = xb + (long)c * Tn * HW;
                const float* wc = w1 + ((long)c2 * C + c) * k;
                for (int dt = 0; dt < k; ++dt) {
                    const float wv = wc[dt];
#pragma unroll
                    for (int j = 0; j < PT_CHUNK; ++j) {
                        const int tt = t0 + j + dt;
                        if (t0 + j < Tp) acc[j] = fmaf(wv, xc[(long)tt * HW], acc[j]);
                    }
                }
            }
#pragma unroll
            for (int j = 0; j < PT_CHUNK; ++j)
                if (t0 + j < Tp) cnb_st(up + t0 + j, acc[j]);
        }
    }
}

constexpr int PT_MAX_K = 8;

// dw1[c2][c][dt] += sum_{p,t'} du[p][c2*T'+t'] * x[b,c,t'+dt,hw] ; grid.y = C*C pairs
template <typename T>
__global__ void __launch_bounds__(256) pretime_conv_wgrad_kernel(const float* __restrict__ x, const T* __restrict__ du,
                                                                float* __restrict__ dw1, int B, int C, int Tn, int H, int W, int k) {
    __shared__ float red[PT_MAX_K][8];
    const int Tp = Tn - k + 1;
    const long HW = (long)H * W;
    const long total = (long)B * HW;
    const int c2 = blockIdx.y / C, c = blockIdx.y - c2 * C;
    float acc[PT_MAX_K];
#pragma unroll
    for (int j = 0; j < PT_MAX_K; ++j) acc[j] = 0.f;
    for (long p = (long)blockIdx.x * blockDim.x + threadIdx.x; p < total; p += (long)gridDim.x * blockDim.x) {
        const long b = p / HW, hw = p - b * HW;
        const float* xc = x + (b * C + c) * Tn * HW + hw;
        const T* dup = du + p * ((long)C * Tp) + (long)c2 * Tp;
        for (int t = 0; t < Tp; ++t) {
            const float g = cnb_ld(dup + t);
#pragma unroll
            for (int dt = 0; dt < PT_MAX_K; ++dt)
                if (dt < k) acc[dt] = fmaf(g, xc[(long)(t + dt) * HW], acc[dt]);
        }
    }
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (int dt = 0; dt < PT_MAX_K; ++dt) {
        const float s = cnb_warp_sum(acc[dt]);
        if (lane == 0) red[dt][wid] = s;
    }
    __syncthreads();
    if (threadIdx.x < k) {
        float s = 0.f;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) s += red[threadIdx.x][w];
        atomicAdd(dw1 + ((long)c2 * C + c) * k + threadIdx.x, s);
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
