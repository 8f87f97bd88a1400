// 128-bit versions of the bandwidth-bound kernels (BatchNorm statistics / apply / backward, n-ary add, LayerNorm, bilinear
// resize, bias gradient) for pixel-major [P][C] tensors whose channel count is a multiple of the 16-byte vector width
// (8 x bf16 or 4 x fp32).  Every thread moves whole 16-byte vectors and stays on ONE channel group for its whole life
// (the grid stride is a multiple of the row length), so per-channel parameters live in registers and there is no
// integer division in the streaming loop.  cnb_api.cu picks these when alignment and divisibility allow and the scalar
// kernels of k_norm.cuh / k_misc.cuh otherwise.
#pragma once
#include "cnb_common.cuh"
#include "k_misc.cuh"
#include "k_norm.cuh"

namespace cnb {

// ---------------------------------------------------------------------------------------------------------------------
// BatchNorm: x is [P][C], vector index i covers channels (i % CV)*V .. +V
// ---------------------------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) bn_stats_vec_kernel(const T* __restrict__ x, long total_v, int CV, long stride_v, int C,
                                                          int ch_div, float* __restrict__ sums) {
    CNB_PDL_SYNC();
    constexpr int V = cnb_vec<T>::N;
    CNB_DYN_SMEM(sh_raw);  // 2*C floats
    float* sh_dyn = reinterpret_cast<float*>(sh_raw);
    for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) sh_dyn[i] = 0.f;
    __syncthreads();
    const long gid = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid < stride_v) {
        const int c0 = (int)(gid % CV) * V;
        float s[V], q[V];
#pragma unroll
        for (int j = 0; j < V; ++j) s[j] = 0.f, q[j] = 0.f;
        // U independent 16-byte loads are issued before any of them is consumed: bytes in flight, not occupancy, feed HBM
        constexpr int U = 4;
        for (long i = gid; i < total_v; i += U * stride_v) {
            float v[U][V];
#pragma unroll
            for (int u = 0; u < U; ++u)
                if (i + u * stride_v < total_v) cnb_ldv(x + (i + u * stride_v) * V, v[u]);
#pragma unroll
            for (int u = 0; u < U; ++u)
                if (i + u * stride_v < total_v) {
#pragma unroll
                    for (int j = 0; j < V; ++j) {
                        s[j] += v[u][j];
                        q[j] = fmaf(v[u][j], v[u][j], q[j]);
                    }
                }
        }
#pragma unroll
        for (int j = 0; j < V; ++j) {
            const int ch = (c0 + j) / ch_div;  // columns past C*ch_div are row padding
            if (ch < C) {
                atomicAdd(&sh_dyn[ch], s[j]);
                atomicAdd(&sh_dyn[C + ch], q[j]);
            }
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) atomicAdd(&sums[i], sh_dyn[i]);
}

template <typename T>
__global__ void __launch_bounds__(256) bn_act_fwd_vec_kernel(const T* __restrict__ x, const float* __restrict__ scale,
                                                            const float* __restrict__ shift, const T* __restrict__ residual,
                                                            T* __restrict__ y, long total_v, int CV, long stride_v, int C, int ch_div,
                                                            int act) {
    CNB_PDL_SYNC();
    constexpr int V = cnb_vec<T>::N;
    const long gid = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= stride_v) return;
    const int c0 = (int)(gid % CV) * V;
    float sc[V], sf[V];
#pragma unroll
    for (int j = 0; j < V; ++j) {
        const int ch = (c0 + j) / ch_div;
        sc[j] = ch < C ? scale[ch] : 0.f;  // padding columns come out as act(0) = 0
        sf[j] = ch < C ? shift[ch] : 0.f;
    }
    constexpr int U = 2;
    for (long i = gid; i < total_v; i += U * stride_v) {
        float v[U][V], r[U][V];
#pragma unroll
        for (int u = 0; u < U; ++u)
            if (i + u * stride_v < total_v) {
                cnb_ldv(x + (i + u * stride_v) * V, v[u]);
                if (residual) cnb_ldv(residual + (i + u * stride_v) * V, r[u]);
            }
#pragma unroll
        for (int u = 0; u < U; ++u)
            if (i + u * stride_v < total_v) {
#pragma unroll
                for (int j = 0; j < V; ++j) {
                    float z = fmaf(v[u][j], sc[j], sf[j]);
                    z = cnb_act_t<T>(z, act);
                    v[u][j] = residual ? z + r[u][j] : z;
                }
                cnb_stv(y + (i + u * stride_v) * V, v[u]);
            }
    }
}

template <typename T>
__global__ void __launch_bounds__(256, 3) bn_act_bwd_reduce_vec_kernel(const T* __restrict__ x, const T* __restrict__ dy,
                                                                   const float* __restrict__ mean, const float* __restrict__ rstd,
                                                                   const float* __restrict__ gamma, const float* __restrict__ beta,
                                                                   long total_v, int CV, long stride_v, int C, int ch_div, int act,
                                                                   float* __restrict__ dsums) {
    CNB_PDL_SYNC();
    constexpr int V = cnb_vec<T>::N;
    CNB_DYN_SMEM(sh_raw);
    float* sh_dyn = reinterpret_cast<float*>(sh_raw);
    for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) sh_dyn[i] = 0.f;
    __syncthreads();
    const long gid = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid < stride_v) {
        const int c0 = (int)(gid % CV) * V;
        // z = x*A + Bc with A = gamma*rstd, Bc = beta - mean*A (the forward's scale/shift); xhat = (x - mean)*rstd is applied to the
        // sum at the end, so only three per-channel constants stay in registers
        float mu[V], A[V], Bc[V], s[V], sx[V];
#pragma unroll
        int chn[V];
        for (int j = 0; j < V; ++j) {
            const int ch = (c0 + j) / ch_div;
            chn[j] = ch < C ? ch : -1;
            const int cc = ch < C ? ch : 0;
            const float gj = gamma ? gamma[cc] : 1.f, bj = beta ? beta[cc] : 0.f;
            mu[j] = mean[cc];
            A[j] = gj * rstd[cc];
            Bc[j] = bj - mu[j] * A[j];
            s[j] = 0.f, sx[j] = 0.f;
        }
        constexpr int U = 2;
        for (long i = gid; i < total_v; i += U * stride_v) {
            float xv[U][V], dv[U][V];
#pragma unroll
            for (int u = 0; u < U; ++u)
                if (i + u * stride_v < total_v) {
                    cnb_ldv(x + (i + u * stride_v) * V, xv[u]);
                    cnb_ldv(dy + (i + u * stride_v) * V, dv[u]);
                }
#pragma unroll
            for (int u = 0; u < U; ++u)
                if (i + u * stride_v < total_v) {
#pragma unroll
                    for (int j = 0; j < V; ++j) {
                        float dz = dv[u][j];
                        if (act) dz *= cnb_act_grad_t<T>(fmaf(xv[u][j], A[j], Bc[j]), act);
                        s[j] += dz;
                        sx[j] = fmaf(dz, xv[u][j] - mu[j], sx[j]);
                    }
                }
        }
#pragma unroll
        for (int j = 0; j < V; ++j) {
            if (chn[j] >= 0) {
                atomicAdd(&sh_dyn[chn[j]], s[j]);
                atomicAdd(&sh_dyn[C + chn[j]], sx[j] * rstd[chn[j]]);
            }
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) atomicAdd(&dsums[i], sh_dyn[i]);
}

template <typename T>
__global__ void __launch_bounds__(256, 3) bn_act_bwd_apply_vec_kernel(const T* __restrict__ x, const T* __restrict__ dy,
                                                                  const float* __restrict__ mean, const float* __restrict__ rstd,
                                                                  const float* __restrict__ gamma, const float* __restrict__ beta,
                                                                  const float* __restrict__ dsums, float inv_count, T* __restrict__ dx,
                                                                  long total_v, int CV, long stride_v, int C, int ch_div, int act,
                                                                  int train_stats) {
    CNB_PDL_SYNC();
    constexpr int V = cnb_vec<T>::N;
    const long gid = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= stride_v) return;
    const int c0 = (int)(gid % CV) * V;
    // dx = A*dz - x*K1 + Q with A = gamma*rstd, K1 = A*rstd*mean(dz*xhat), Q = mean*K1 - A*mean(dz); z = x*A + Bc
    float A[V], Bc[V], K1[V], Q[V];
#pragma unroll
    for (int j = 0; j < V; ++j) {
        const int ch = (c0 + j) / ch_div;
        if (ch < C) {
            const float gj = gamma ? gamma[ch] : 1.f, bj = beta ? beta[ch] : 0.f;
            const float m = mean[ch], r = rstd[ch];
            A[j] = gj * r;
            Bc[j] = bj - m * A[j];
            K1[j] = train_stats ? A[j] * r * dsums[C + ch] * inv_count : 0.f;
            Q[j] = train_stats ? m * K1[j] - A[j] * dsums[ch] * inv_count : 0.f;
        } else {  // row padding: gradient zero (dy there must be finite)
            A[j] = 0.f, Bc[j] = 0.f, K1[j] = 0.f, Q[j] = 0.f;
        }
    }
    constexpr int U = 2;
    for (long i = gid; i < total_v; i += U * stride_v) {
        float xv[U][V], dv[U][V];
#pragma unroll
        for (int u = 0; u < U; ++u)
            if (i + u * stride_v < total_v) {
                cnb_ldv(x + (i + u * stride_v) * V, xv[u]);
                cnb_ldv(dy + (i + u * stride_v) * V, dv[u]);
            }
#pragma unroll
        for (int u = 0; u < U; ++u)
            if (i + u * stride_v < total_v) {
#pragma unroll
                for (int j = 0; j < V; ++j) {
                    float dz = dv[u][j];
                    if (act) dz *= cnb_act_grad_t<T>(fmaf(xv[u][j], A[j], Bc[j]), act);
                    xv[u][j] = fmaf(A[j], dz, fmaf(-xv[u][j], K1[j], Q[j]));
                }
                cnb_stv(dx + (i + u * stride_v) * V, xv[u]);
            }
    }
}

template <typename T>
__global__ void __launch_bounds__(256) add_n_vec_kernel(const T* __restrict__ a, const T* __restrict__ b, const T* __restrict__ c,
                                                       const T* __restrict__ d, T* __restrict__ out, long n_v) {
    CNB_PDL_SYNC();
    constexpr int V = cnb_vec<T>::N;
#pragma unroll 2
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n_v; i += (long)gridDim.x * blockDim.x) {
        float v[V], w[V];
        cnb_ldv(a + i * V, v);
        cnb_ldv(b + i * V, w);
#pragma unroll
        for (int j = 0; j < V; ++j) v[j] += w[j];
        if (c) {
            cnb_ldv(c + i * V, w);
#pragma unroll
            for (int j = 0; j < V; ++j) v[j] += w[j];
        }
        if (d) {
            cnb_ldv(d + i * V, w);
#pragma unroll
            for (int j = 0; j < V; ++j) v[j] += w[j];
        }
        cnb_stv(out + i * V, v);
    }
}

// db[n] = sum_p dy[p][n] for a dense [P][N] gradient (N % V == 0)
template <typename T>
__global__ void __launch_bounds__(256) colsum_vec_kernel(const T* __restrict__ dy, long total_v, int CV, long stride_v, int N,
                                                        float* __restrict__ db) {
    CNB_PDL_SYNC();
    constexpr int V = cnb_vec<T>::N;
    CNB_DYN_SMEM(sh_raw);
    float* sh_dyn = reinterpret_cast<float*>(sh_raw);
    for (int i = threadIdx.x; i < N; i += blockDim.x) sh_dyn[i] = 0.f;
    __syncthreads();
    const long gid = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid < stride_v) {
        const int c0 = (int)(gid % CV) * V;
        float s[V];
#pragma unroll
        for (int j = 0; j < V; ++j) s[j] = 0.f;
        constexpr int U = 4;
        for (long i = gid; i < total_v; i += U * stride_v) {
            float v[U][V];
#pragma unroll
            for (int u = 0; u < U; ++u)
                if (i + u * stride_v < total_v) cnb_ldv(dy + (i + u * stride_v) * V, v[u]);
#pragma unroll
            for (int u = 0; u < U; ++u)
                if (i + u * stride_v < total_v) {
#pragma unroll
                    for (int j = 0; j < V; ++j) s[j] += v[u][j];
                }
        }
#pragma unroll
        for (int j = 0; j < V; ++j) atomicAdd(&sh_dyn[c0 + j], s[j]);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < N; i += blockDim.x) atomicAdd(&db[i], sh_dyn[i]);
}

// ---------------------------------------------------------------------------------------------------------------------
// LayerNorm over channels: one warp per pixel, K vectors per lane (C <= 32*V*K), one read of x
// ---------------------------------------------------------------------------------------------------------------------
template <typename T, int K>
__global__ void __launch_bounds__(256, 3) layernorm_fwd_vec_kernel(const T* __restrict__ x, const float* __restrict__ gamma,
                                                               const float* __restrict__ beta, float eps, T* __restrict__ y,
                                                               float* __restrict__ save_mean, float* __restrict__ save_rstd, long P,
                                                               int C) {
    CNB_PDL_SYNC();
    constexpr int V = cnb_vec<T>::N;
    const int lane = threadIdx.x & 31;
    const long warp = ((long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long nwarps = ((long)gridDim.x * blockDim.x) >> 5;
    const float invC = 1.0f / (float)C;
    float gm[K][V], bt[K][V];  // this lane's affine parameters (ncu: re-reading them per pixel kept L1 95 % busy)
#pragma unroll
    for (int k = 0; k < K; ++k) {
        const int c = (lane + 32 * k) * V;
#pragma unroll
        for (int j = 0; j < V; ++j) {
            gm[k][j] = c < C ? gamma[c + j] : 0.f;
            bt[k][j] = c < C ? beta[c + j] : 0.f;
        }
    }
    // U pixels per warp and trip, all loads issued before the first reduction: a warp moves 512 bytes per pixel at C = 256, and with one
    // pixel in flight per warp an SM had ~16 KB outstanding (ncu: 4.0 TB/s, issue slots 49 % busy, waiting on the loads)
    constexpr int U = K == 1 ? 4 : 2;
    for (long p0 = warp * U; p0 < P; p0 += nwarps * U) {
        uint4 raw[U][K];  // in flight as raw 16-byte words: 4 registers per vector instead of 8 expanded floats
#pragma unroll
        for (int u = 0; u < U; ++u)
#pragma unroll
            for (int k = 0; k < K; ++k) {
                const int c = (lane + 32 * k) * V;
                if (c < C && p0 + u < P) raw[u][k] = cnb_ldraw(x + (p0 + u) * C + c);
            }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const long p = p0 + u;
            if (p >= P) break;  // warp-uniform
            float v[1][K][V];
            float s = 0.f;
#pragma unroll
            for (int k = 0; k < K; ++k) {
                const int c = (lane + 32 * k) * V;
                if (c < C) {
                    cnb_expand(raw[u][k], v[0][k], x);
#pragma unroll
                    for (int j = 0; j < V; ++j) s += v[0][k][j];
                }
            }
            const float mean = cnb_warp_sum(s) * invC;
            float q = 0.f;
#pragma unroll
            for (int k = 0; k < K; ++k) {
                const int c = (lane + 32 * k) * V;
                if (c < C) {
#pragma unroll
                    for (int j = 0; j < V; ++j) {
                        const float d = v[0][k][j] - mean;
                        q = fmaf(d, d, q);
                    }
                }
            }
            const float rstd = rsqrtf(cnb_warp_sum(q) * invC + eps);
#pragma unroll
            for (int k = 0; k < K; ++k) {
                const int c = (lane + 32 * k) * V;
                if (c < C) {
#pragma unroll
                    for (int j = 0; j < V; ++j) v[0][k][j] = fmaf((v[0][k][j] - mean) * rstd, gm[k][j], bt[k][j]);
                    cnb_stv(y + p * C + c, v[0][k]);
                }
            }
            if (lane == 0) {
                save_mean[p] = mean;
                save_rstd[p] = rstd;
            }
        }
    }
}

template <typename T, int K>
__global__ void __launch_bounds__(256, 3) layernorm_bwd_vec_kernel(const T* __restrict__ x, const T* __restrict__ dy,
                                                               const float* __restrict__ gamma, const float* __restrict__ save_mean,
                                                               const float* __restrict__ save_rstd, T* __restrict__ dx,
                                                               float* __restrict__ dgamma, float* __restrict__ dbeta, long P, int C) {
    CNB_PDL_SYNC();
    constexpr int V = cnb_vec<T>::N;
    CNB_DYN_SMEM(sh_raw);  // 2*C floats
    float* sh_dyn = reinterpret_cast<float*>(sh_raw);
    for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) sh_dyn[i] = 0.f;
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const long warp = ((long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long nwarps = ((long)gridDim.x * blockDim.x) >> 5;
    const float invC = 1.0f / (float)C;
    float g[K][V], dg[K][V], db[K][V];
#pragma unroll
    for (int k = 0; k < K; ++k) {
        const int c = (lane + 32 * k) * V;
#pragma unroll
        for (int j = 0; j < V; ++j) {
            g[k][j] = c < C ? gamma[c + j] : 0.f;
            dg[k][j] = 0.f;
            db[k][j] = 0.f;
        }
    }
    constexpr int U = K == 1 ? 2 : 1;  // pixels per warp and trip (see the forward kernel)
    for (long p0 = warp * U; p0 < P; p0 += nwarps * U) {
        uint4 rx[U][K], rd[U][K];
        float mean[U], rstd[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const long p = p0 + u < P ? p0 + u : P - 1;
            mean[u] = save_mean[p], rstd[u] = save_rstd[p];
#pragma unroll
            for (int k = 0; k < K; ++k) {
                const int c = (lane + 32 * k) * V;
                if (c < C && p0 + u < P) {
                    rx[u][k] = cnb_ldraw(x + p * C + c);
                    rd[u][k] = cnb_ldraw(dy + p * C + c);
                }
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const long p = p0 + u;
            if (p >= P) break;  // warp-uniform
            float xh[1][K][V], d[1][K][V];
            float s1 = 0.f, s2 = 0.f;
#pragma unroll
            for (int k = 0; k < K; ++k) {
                const int c = (lane + 32 * k) * V;
                if (c < C) {
                    cnb_expand(rx[u][k], xh[0][k], x);
                    cnb_expand(rd[u][k], d[0][k], dy);
#pragma unroll
                    for (int j = 0; j < V; ++j) {
                        xh[0][k][j] = (xh[0][k][j] - mean[u]) * rstd[u];
                        const float gd = d[0][k][j] * g[k][j];
                        s1 += gd;
                        s2 = fmaf(gd, xh[0][k][j], s2);
                    }
                }
            }
            s1 = cnb_warp_sum(s1) * invC;
            s2 = cnb_warp_sum(s2) * invC;
#pragma unroll
            for (int k = 0; k < K; ++k) {
                const int c = (lane + 32 * k) * V;
                if (c < C) {
                    float o[V];
#pragma unroll
                    for (int j = 0; j < V; ++j) {
                        o[j] = rstd[u] * (d[0][k][j] * g[k][j] - s1 - xh[0][k][j] * s2);
                        dg[k][j] = fmaf(d[0][k][j], xh[0][k][j], dg[k][j]);
                        db[k][j] += d[0][k][j];
                    }
                    cnb_stv(dx + p * C + c, o);
                }
            }
        }
    }
#pragma unroll
    for (int k = 0; k < K; ++k) {
        const int c = (lane + 32 * k) * V;
        if (c < C) {
#pragma unroll
            for (int j = 0; j < V; ++j) {
                atomicAdd(&sh_dyn[c + j], dg[k][j]);
                atomicAdd(&sh_dyn[C + c + j], db[k][j]);
            }
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < C; i += blockDim.x) {
        atomicAdd(dgamma + i, sh_dyn[i]);
        atomicAdd(dbeta + i, sh_dyn[C + i]);
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// bilinear align_corners=True resize.  One CTA per output (forward) / input (backward) image row: the row's vertical taps and
// weights are computed once per CTA, threads walk (x, channel vector) with 32-bit index math only (ncu: the flat-index version
// spent 72 % of its issue slots on 64-bit div/mod and was instruction-bound at 1.3-3.1 TB/s).
// ---------------------------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) resize_bilinear_fwd_vec_kernel(const T* __restrict__ x, T* __restrict__ y, int B, int Hin, int Win,
                                                                     int Hout, int Wout, int C, float rh, float rw) {
    CNB_PDL_SYNC();
    constexpr int V = cnb_vec<T>::N;
    const int CV = C / V;
    const int oy = blockIdx.x % Hout;
    const int b = blockIdx.x / Hout;
    int y0, y1;
    float ly0, ly1;
    bilinear_src(oy, rh, Hin, y0, y1, ly0, ly1);
    const T* row0 = x + ((long)b * Hin + y0) * Win * C;
    const T* row1 = x + ((long)b * Hin + y1) * Win * C;
    T* orow = y + ((long)b * Hout + oy) * Wout * C;
    const int n = Wout * CV;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const int ox = i / CV;
        const int c = (i - ox * CV) * V;
        int x0, x1;
        float lx0, lx1;
        bilinear_src(ox, rw, Win, x0, x1, lx0, lx1);
        float v00[V], v01[V], v10[V], v11[V];
        cnb_ldv(row0 + x0 * C + c, v00);
        cnb_ldv(row0 + x1 * C + c, v01);
        cnb_ldv(row1 + x0 * C + c, v10);
        cnb_ldv(row1 + x1 * C + c, v11);
#pragma unroll
        for (int j = 0; j < V; ++j) v00[j] = ly0 * (lx0 * v00[j] + lx1 * v01[j]) + ly1 * (lx0 * v10[j] + lx1 * v11[j]);
        cnb_stv(orow + ox * C + c, v00);
    }
}

// gather form of the adjoint (deterministic, no atomics): input pixel (iy, ix) collects from the output pixels that read it
template <typename T>
__global__ void __launch_bounds__(256) resize_bilinear_bwd_vec_kernel(const T* __restrict__ dy, T* __restrict__ dx, int B, int Hin,
                                                                     int Win, int Hout, int Wout, int C, float rh, float rw) {
    CNB_PDL_SYNC();
    constexpr int V = cnb_vec<T>::N;
    constexpr int MAXC = 6;  // candidate output rows / columns handled with per-axis weight tables
    const int CV = C / V;
    const int iy = blockIdx.x % Hin;
    const int b = blockIdx.x / Hin;
    int ylo, yhi;
    bilinear_candidates(iy, rh, Hout, ylo, yhi);
    const bool y_table = yhi - ylo < MAXC;
    float wy[MAXC];
#pragma unroll
    for (int a = 0; a < MAXC; ++a) wy[a] = (y_table && ylo + a <= yhi) ? bilinear_weight(ylo + a, iy, rh, Hin) : 0.f;
    const T* base = dy + (long)b * Hout * Wout * C;
    T* orow = dx + ((long)b * Hin + iy) * Win * C;
    const int n = Win * CV;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const int ix = i / CV;
        const int c = (i - ix * CV) * V;
        int xlo, xhi;
        bilinear_candidates(ix, rw, Wout, xlo, xhi);
        float acc[V];
#pragma unroll
        for (int j = 0; j < V; ++j) acc[j] = 0.f;
        if (y_table && xhi - xlo < MAXC) {
            float wx[MAXC];
#pragma unroll
            for (int a = 0; a < MAXC; ++a) wx[a] = (xlo + a <= xhi) ? bilinear_weight(xlo + a, ix, rw, Win) : 0.f;
#pragma unroll
            for (int a = 0; a < MAXC; ++a) {
                if (wy[a] == 0.f) continue;
#pragma unroll
                for (int bb = 0; bb < MAXC; ++bb) {
                    const float w = wy[a] * wx[bb];
                    if (w != 0.f) {
                        float v[V];
                        cnb_ldv(base + ((long)(ylo + a) * Wout + (xlo + bb)) * C + c, v);
#pragma unroll
                        for (int j = 0; j < V; ++j) acc[j] = fmaf(w, v[j], acc[j]);
                    }
                }
            }
        } else {
            for (int oy = ylo; oy <= yhi; ++oy) {
                const float wyy = bilinear_weight(oy, iy, rh, Hin);
                if (wyy == 0.f) continue;
                for (int ox = xlo; ox <= xhi; ++ox) {
                    const float w = wyy * bilinear_weight(ox, ix, rw, Win);
                    if (w != 0.f) {
                        float v[V];
                        cnb_ldv(base + ((long)oy * Wout + ox) * C + c, v);
#pragma unroll
                        for (int j = 0; j < V; ++j) acc[j] = fmaf(w, v[j], acc[j]);
                    }
                }
            }
        }
        cnb_stv(orow + ix * C + c, acc);
    }
}

// Table form of the adjoint for scale factors >= 0.5 (the ConvTranspose fix-up: 127 -> 128), where an input index is read by at most
// RB_NC consecutive output indices per axis.  One CTA per input row: the x-axis supports (first output column + RB_NC weights per input
// column) are computed once per CTA into shared memory, the y-axis support is CTA-uniform, and the streaming loop is
// `table lookup + at most RB_NC^2 guarded 16-byte loads` (ncu on the generic gather kernel: 74 % issue-slot use, 2.0 TB/s, most of
// it weight arithmetic for candidates whose weight is zero).
constexpr int RB_NC = 5;       // scale >= 0.5
constexpr int RB_NC_NEAR1 = 3; // scale >= 0.7: an input index is read by at most 3 outputs (support of length 2/scale < 3): the
                               // ConvTranspose fix-up (2H-1 -> 2H) walks 9 candidates per vector instead of 25
// tight support of input index i along one axis: first output index with a non-zero weight, and RB_NC weights from there
template <int NC>
__device__ __forceinline__ void bilinear_support(int i, float rscale, int in_len, int out_len, int& first, float* w) {
    int lo, hi;
    bilinear_candidates(i, rscale, out_len, lo, hi);
    first = lo;
    bool found = false;
#pragma unroll
    for (int a = 0; a < NC; ++a) w[a] = 0.f;
    for (int o = lo; o <= hi; ++o) {
        const float wo = bilinear_weight(o, i, rscale, in_len);
        if (!found && wo != 0.f) {
            found = true;
            first = o;
        }
        if (found && o - first < NC) w[o - first] = wo;
    }
}

// One CTA per input row (the x-axis table is rebuilt by every CTA, but 4064 short CTAs fill and drain the machine better than ~1000
// persistent ones).  COLSUM is a template parameter: the eight extra accumulators cost a resident CTA per SM (56 registers: four CTAs
// instead of six), which this latency-bound kernel pays for in full (ncu: 159 -> 212 us at level a when the plain variant carried them).  With `colsum` the CTA also adds the per-channel sums of the row it wrote to colsum[C]
// (the bias gradient of the ConvTranspose2d whose output was resized).
template <typename T, int NC, bool COLSUM>
__global__ void __launch_bounds__(256, COLSUM ? 5 : 6) resize_bilinear_bwd_row_kernel(const T* __restrict__ dy, T* __restrict__ dx, int B, int Hin, int Win,
                                                                     int Hout, int Wout, int C, float rh, float rw,
                                                                     float* __restrict__ colsum) {
    CNB_PDL_SYNC();
    constexpr int V = cnb_vec<T>::N;
    CNB_DYN_SMEM(sm_raw);  // int xfirst[Win]; float wx[Win][NC]; int yfirst; float wy[NC]
    int* xfirst = reinterpret_cast<int*>(sm_raw);
    float* wxs = reinterpret_cast<float*>(xfirst + Win);
    int* yfirst_s = reinterpret_cast<int*>(wxs + Win * NC);
    float* wys = reinterpret_cast<float*>(yfirst_s + 1);
    float* csum = wys + NC;  // [C] when colsum
    const int CV = C / V;
    const int iy = blockIdx.x % Hin;
    const int b = blockIdx.x / Hin;
    for (int ix = threadIdx.x; ix < Win; ix += blockDim.x) {
        float w[NC];
        int f;
        bilinear_support<NC>(ix, rw, Win, Wout, f, w);
        xfirst[ix] = f;
#pragma unroll
        for (int a = 0; a < NC; ++a) wxs[ix * NC + a] = w[a];
    }
    if (COLSUM)
        for (int i = threadIdx.x; i < C; i += blockDim.x) csum[i] = 0.f;
    if (threadIdx.x == 0) {
        float w[NC];
        int f;
        bilinear_support<NC>(iy, rh, Hin, Hout, f, w);
        *yfirst_s = f;
#pragma unroll
        for (int a = 0; a < NC; ++a) wys[a] = w[a];
    }
    __syncthreads();
    const int yfirst = *yfirst_s;
    float wy[NC];
#pragma unroll
    for (int a = 0; a < NC; ++a) wy[a] = wys[a];
    const T* base = dy + ((long)b * Hout + yfirst) * Wout * C;
    T* orow = dx + ((long)b * Hin + iy) * Win * C;
    const int n = Win * CV;
    float cs[V];
#pragma unroll
    for (int j = 0; j < V; ++j) cs[j] = 0.f;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const int ix = i / CV;
        const int c = (i - ix * CV) * V;
        const T* col = base + (long)xfirst[ix] * C + c;
        float wx[NC];
#pragma unroll
        for (int a = 0; a < NC; ++a) wx[a] = wxs[ix * NC + a];
        float acc[V];
#pragma unroll
        for (int j = 0; j < V; ++j) acc[j] = 0.f;
        if (NC == 3 && wy[2] == 0.f && wx[2] == 0.f) {
            // the 2H-1 -> 2H fix-up: exactly two outputs read an input index per axis (a third only through a rounding of the last
            // output's source coordinate).  Four unconditional loads, issued together, instead of nine guarded candidates: the same
            // shape as the forward kernel, which streams at 0.73 of the copy rate where this loop reached 0.5.
            const long r1 = wy[1] != 0.f ? (long)Wout * C : 0, c1 = wx[1] != 0.f ? C : 0;  // stay inside the tensor when the weight is 0
            const uint4 q00 = cnb_ldraw(col), q01 = cnb_ldraw(col + c1), q10 = cnb_ldraw(col + r1), q11 = cnb_ldraw(col + r1 + c1);
            const float w00 = wy[0] * wx[0], w01 = wy[0] * wx[1], w10 = wy[1] * wx[0], w11 = wy[1] * wx[1];
            float v[V];
            cnb_expand(q00, v, col);
#pragma unroll
            for (int j = 0; j < V; ++j) acc[j] = w00 * v[j];
            cnb_expand(q01, v, col);
#pragma unroll
            for (int j = 0; j < V; ++j) acc[j] = fmaf(w01, v[j], acc[j]);
            cnb_expand(q10, v, col);
#pragma unroll
            for (int j = 0; j < V; ++j) acc[j] = fmaf(w10, v[j], acc[j]);
            cnb_expand(q11, v, col);
#pragma unroll
            for (int j = 0; j < V; ++j) acc[j] = fmaf(w11, v[j], acc[j]);
        } else {
#pragma unroll
            for (int a = 0; a < NC; ++a) {
                if (wy[a] == 0.f) continue;  // CTA-uniform
#pragma unroll
                for (int bb = 0; bb < NC; ++bb) {
                    const float w = wy[a] * wx[bb];
                    if (w != 0.f) {
                        float v[V];
                        cnb_ldv(col + ((long)a * Wout + bb) * C, v);
#pragma unroll
                        for (int j = 0; j < V; ++j) acc[j] = fmaf(w, v[j], acc[j]);
                    }
                }
            }
        }
        cnb_stv(orow + ix * C + c, acc);
        if (COLSUM) {
#pragma unroll
            for (int j = 0; j < V; ++j) cs[j] += acc[j];
        }
    }
    if (COLSUM) {  // blockDim % CV == 0: a thread stays on one channel vector
        if ((int)threadIdx.x < n) {
            const int c = ((int)threadIdx.x % CV) * V;
#pragma unroll
            for (int j = 0; j < V; ++j) atomicAdd(&csum[c + j], cs[j]);
        }
        __syncthreads();
        for (int i = threadIdx.x; i < C; i += blockDim.x) atomicAdd(&colsum[i], csum[i]);
    }
}


// Persistent over input rows (row = blockIdx.x, += gridDim.x): the x- and y-axis support tables are built ONCE per CTA (the one-row
// CTAs of the first version rebuilt the x table 4064 times per level-a launch), and with `colsum` the kernel also accumulates the
// per-channel sum of the gradient it writes -- the bias gradient of the ConvTranspose2d whose output was resized -- so that tensor is
// not read a second time by a column-sum launch (blockDim % (C / V) == 0: a thread stays on one channel vector).
template <typename T, int NC>
__global__ void __launch_bounds__(256) resize_bilinear_bwd_tab_kernel(const T* __restrict__ dy, T* __restrict__ dx, int B, int Hin, int Win,
                                                                     int Hout, int Wout, int C, float rh, float rw,
                                                                     float* __restrict__ colsum) {
    CNB_PDL_SYNC();
    constexpr int V = cnb_vec<T>::N;
    CNB_DYN_SMEM(sm_raw);  // int xfirst[Win]; float wx[Win][NC]; int yfirst[Hin]; float wy[Hin][NC]; float csum[C]
    int* xfirst = reinterpret_cast<int*>(sm_raw);
    float* wxs = reinterpret_cast<float*>(xfirst + Win);
    int* yfirst_s = reinterpret_cast<int*>(wxs + Win * NC);
    float* wys = reinterpret_cast<float*>(yfirst_s + Hin);
    float* csum = wys + Hin * NC;
    const int CV = C / V;
    for (int ix = threadIdx.x; ix < Win; ix += blockDim.x) {
        float w[NC];
        int f;
        bilinear_support<NC>(ix, rw, Win, Wout, f, w);
        xfirst[ix] = f;
#pragma unroll
        for (int a = 0; a < NC; ++a) wxs[ix * NC + a] = w[a];
    }
    for (int iy = threadIdx.x; iy < Hin; iy += blockDim.x) {
        float w[NC];
        int f;
        bilinear_support<NC>(iy, rh, Hin, Hout, f, w);
        yfirst_s[iy] = f;
#pragma unroll
        for (int a = 0; a < NC; ++a) wys[iy * NC + a] = w[a];
    }
    if (colsum)
        for (int i = threadIdx.x; i < C; i += blockDim.x) csum[i] = 0.f;
    __syncthreads();
    float cs[V];
#pragma unroll
    for (int j = 0; j < V; ++j) cs[j] = 0.f;
    const int n = Win * CV;
    for (int row = blockIdx.x; row < B * Hin; row += gridDim.x) {
        const int iy = row % Hin;
        const int b = row / Hin;
        float wy[NC];
#pragma unroll
        for (int a = 0; a < NC; ++a) wy[a] = wys[iy * NC + a];
        const T* base = dy + ((long)b * Hout + yfirst_s[iy]) * Wout * C;
        T* orow = dx + ((long)b * Hin + iy) * Win * C;
        for (int i = threadIdx.x; i < n; i += blockDim.x) {
            const int ix = i / CV;
            const int c = (i - ix * CV) * V;
            const T* col = base + (long)xfirst[ix] * C + c;
            float wx[NC];
#pragma unroll
            for (int a = 0; a < NC; ++a) wx[a] = wxs[ix * NC + a];
            float acc[V];
#pragma unroll
            for (int j = 0; j < V; ++j) acc[j] = 0.f;
            if (NC == 3 && wy[2] == 0.f && wx[2] == 0.f) {  // the 2 x 2 case of the row kernel above
                const long r1 = wy[1] != 0.f ? (long)Wout * C : 0, c1 = wx[1] != 0.f ? C : 0;
                const uint4 q00 = cnb_ldraw(col), q01 = cnb_ldraw(col + c1), q10 = cnb_ldraw(col + r1), q11 = cnb_ldraw(col + r1 + c1);
                const float w00 = wy[0] * wx[0], w01 = wy[0] * wx[1], w10 = wy[1] * wx[0], w11 = wy[1] * wx[1];
                float v[V];
                cnb_expand(q00, v, col);
#pragma unroll
                for (int j = 0; j < V; ++j) acc[j] = w00 * v[j];
                cnb_expand(q01, v, col);
#pragma unroll
                for (int j = 0; j < V; ++j) acc[j] = fmaf(w01, v[j], acc[j]);
                cnb_expand(q10, v, col);
#pragma unroll
                for (int j = 0; j < V; ++j) acc[j] = fmaf(w10, v[j], acc[j]);
                cnb_expand(q11, v, col);
#pragma unroll
                for (int j = 0; j < V; ++j) acc[j] = fmaf(w11, v[j], acc[j]);
            } else {
#pragma unroll
                for (int a = 0; a < NC; ++a) {
                    if (wy[a] == 0.f) continue;  // CTA-uniform
#pragma unroll
                    for (int bb = 0; bb < NC; ++bb) {
                        const float w = wy[a] * wx[bb];
                        if (w != 0.f) {
                            float v[V];
                            cnb_ldv(col + ((long)a * Wout + bb) * C, v);
#pragma unroll
                            for (int j = 0; j < V; ++j) acc[j] = fmaf(w, v[j], acc[j]);
                        }
                    }
                }
            }
            cnb_stv(orow + ix * C + c, acc);
#pragma unroll
            for (int j = 0; j < V; ++j) cs[j] += acc[j];
        }
    }
    if (colsum) {  // uniform
        if ((int)threadIdx.x < n) {
            const int c = ((int)threadIdx.x % CV) * V;
#pragma unroll
            for (int j = 0; j < V; ++j) atomicAdd(&csum[c + j], cs[j]);
        }
        __syncthreads();
        for (int i = threadIdx.x; i < C; i += blockDim.x) atomicAdd(&colsum[i], csum[i]);
    }
}

}  // namespace cnb
