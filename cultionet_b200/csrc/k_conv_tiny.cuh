// Convolutions with a handful of channels on both sides (the Psi-Net heads' 3->1 and 3->3 convolutions,
// reference nn/modules/unet_parts.py:196-309): pure bandwidth / latency work, one thread per output pixel with the whole
// filter bank in shared memory.  Same descriptor and gather rules as the implicit-GEMM kernels (any stride, direct or
// transposed, several sources), fp32 or bf16 storage, fp32 accumulate.
#pragma once
#include "cnb_common.cuh"
#include "k_conv_generic.cuh"

namespace cnb {

constexpr int TINY_MAX_N = 16;    // output channels
constexpr int TINY_MAX_C = 16;    // input channels over all sources
constexpr int TINY_MAX_TAPS = 16;
constexpr int TINY_WG_MAX_C = 4;  // source channels per CTA of the weight-gradient kernel (blockIdx.z walks chunks of 4)
constexpr int TINY_WG_MAX_N = 8;

inline bool conv_tiny_eligible(const cnb_conv_desc* d) {
    int ctot = 0;
    for (int s = 0; s < d->nsrc; ++s) ctot += d->src_c[s];
    return d->N <= TINY_MAX_N && ctot <= TINY_MAX_C && d->KH * d->KW <= TINY_MAX_TAPS;
}

inline bool wgrad_tiny_eligible(const cnb_wgrad_desc* d) {
    return d->N <= TINY_WG_MAX_N && d->src_c <= TINY_MAX_C && d->KH * d->KW <= TINY_MAX_TAPS;
}

template <typename T>
__global__ void __launch_bounds__(256) conv_tiny_kernel(cnb_conv_desc d, int ctot) {
    CNB_PDL_SYNC();
    __shared__ float ws[TINY_MAX_TAPS * TINY_MAX_N * TINY_MAX_C];
    __shared__ float bs[TINY_MAX_N];
    const int taps = d.KH * d.KW;
    const T* wbase = reinterpret_cast<const T*>(d.w_packed);
    for (int i = threadIdx.x; i < taps * d.N * ctot; i += blockDim.x) {
        const int c = i % ctot;
        const int t = i / ctot;
        const int n = t % d.N, tap = t / d.N;
        ws[i] = cnb_ld(wbase + (long)tap * d.w_tap_stride + (long)n * d.w_row_stride + c);
    }
    if (threadIdx.x < TINY_MAX_N) bs[threadIdx.x] = (d.bias && (int)threadIdx.x < d.N) ? d.bias[threadIdx.x] : 0.f;
    __syncthreads();

    const long M = (long)d.B * d.Hout * d.Wout;
    T* out = reinterpret_cast<T*>(d.out);
    for (long m = (long)blockIdx.x * blockDim.x + threadIdx.x; m < M; m += (long)gridDim.x * blockDim.x) {
        const int ox = (int)(m % d.Wout);
        const long t = m / d.Wout;
        const int oy = (int)(t % d.Hout);
        const int ob = (int)(t / d.Hout);
        float acc[TINY_MAX_N];
#pragma unroll
        for (int n = 0; n < TINY_MAX_N; ++n) acc[n] = bs[n];
        for (int tap = 0; tap < taps; ++tap) {
            const int ky = tap / d.KW, kx = tap - ky * d.KW;
            int iy, ix;
            if (!conv_src_coord(oy, ky, d.stride, d.pad, d.dil, d.transposed, d.Hin, iy)) continue;
            if (!conv_src_coord(ox, kx, d.stride, d.pad, d.dil, d.transposed, d.Win, ix)) continue;
            const long pix = ((long)ob * d.Hin + iy) * d.Win + ix;
            const float* wt = ws + tap * d.N * ctot;
            int coff = 0;
            for (int s = 0; s < d.nsrc; ++s) {
                const T* sp = reinterpret_cast<const T*>(d.src[s]) + pix * d.src_stride[s];
                for (int c = 0; c < d.src_c[s]; ++c) {
                    const float x = cnb_ld(sp + c);
#pragma unroll
                    for (int n = 0; n < TINY_MAX_N; ++n)
                        if (n < d.N) acc[n] = fmaf(x, wt[n * ctot + coff + c], acc[n]);
                }
                coff += d.src_c[s];
            }
        }
#pragma unroll
        for (int n = 0; n < TINY_MAX_N; ++n)
            if (n < d.N) cnb_st(out + m * d.out_stride + n, acc[n]);
    }
}

// The Psi-Net heads only ever use (N, C) = (3, 9), (3, 3) and (9, 3) (the block-diagonal stream outputs, the fuse convolution and
// their data gradients).  With both counts compile-time the channel loops unroll without predicates, a tap's weights for one
// input channel are ONE 16-byte (N = 3) or three 16-byte (N = 9) broadcast shared-memory loads instead of N scalar ones, and the
// accumulators are exactly N registers: ~6x fewer issued instructions per pixel than the generic kernel above, which spent its time
// on predicated-off FMAs of a 16-wide accumulator and one LDS per FMA (ncu: 70-150 us per launch for 9.4 MB of input).
template <typename T, int N, int C>
__global__ void __launch_bounds__(256) conv_tiny_fixed_kernel(cnb_conv_desc d) {
    CNB_PDL_SYNC();
    constexpr int NP = (N + 3) / 4 * 4;  // weights of one (tap, c) padded to whole float4
    __shared__ __align__(16) float ws[TINY_MAX_TAPS * C * NP];
    __shared__ float bs[NP];
    const int taps = d.KH * d.KW;
    const T* wbase = reinterpret_cast<const T*>(d.w_packed);
    for (int i = threadIdx.x; i < taps * C * NP; i += blockDim.x) {
        const int n = i % NP;
        const int t = i / NP;
        const int c = t % C, tap = t / C;
        ws[i] = n < N ? cnb_ld(wbase + (long)tap * d.w_tap_stride + (long)n * d.w_row_stride + c) : 0.f;
    }
    if (threadIdx.x < NP) bs[threadIdx.x] = (d.bias && (int)threadIdx.x < N) ? d.bias[threadIdx.x] : 0.f;
    __syncthreads();

    const long M = (long)d.B * d.Hout * d.Wout;
    T* out = reinterpret_cast<T*>(d.out);
    const T* src = reinterpret_cast<const T*>(d.src[0]);
    const int pitch = d.src_stride[0];
    for (long m = (long)blockIdx.x * blockDim.x + threadIdx.x; m < M; m += (long)gridDim.x * blockDim.x) {
        const int ox = (int)(m % d.Wout);
        const long t = m / d.Wout;
        const int oy = (int)(t % d.Hout);
        const int ob = (int)(t / d.Hout);
        float acc[NP];
#pragma unroll
        for (int n = 0; n < NP; ++n) acc[n] = bs[n];
        for (int tap = 0; tap < taps; ++tap) {
            const int ky = tap / d.KW, kx = tap - ky * d.KW;
            int iy, ix;
            if (!conv_src_coord(oy, ky, d.stride, d.pad, d.dil, d.transposed, d.Hin, iy)) continue;
            if (!conv_src_coord(ox, kx, d.stride, d.pad, d.dil, d.transposed, d.Win, ix)) continue;
            const T* sp = src + (((long)ob * d.Hin + iy) * d.Win + ix) * pitch;
            const float4* wt = reinterpret_cast<const float4*>(ws + tap * C * NP);
            float x[C];
#pragma unroll
            for (int c = 0; c < C; ++c) x[c] = cnb_ld(sp + c);
#pragma unroll
            for (int c = 0; c < C; ++c)
#pragma unroll
                for (int q = 0; q < NP / 4; ++q) {
                    const float4 w4 = wt[c * (NP / 4) + q];
                    acc[4 * q + 0] = fmaf(x[c], w4.x, acc[4 * q + 0]);
                    acc[4 * q + 1] = fmaf(x[c], w4.y, acc[4 * q + 1]);
                    acc[4 * q + 2] = fmaf(x[c], w4.z, acc[4 * q + 2]);
                    acc[4 * q + 3] = fmaf(x[c], w4.w, acc[4 * q + 3]);
                }
        }
#pragma unroll
        for (int n = 0; n < N; ++n) cnb_st(out + m * d.out_stride + n, acc[n]);
    }
}

// dWp[tap][n][k_off + c] += sum_p dY[p][n] * X[gather(p, tap)][c]; grid = (pixel blocks, taps, 4-channel chunks of the source)
template <typename T>
__global__ void __launch_bounds__(256) conv_tiny_wgrad_kernel(cnb_wgrad_desc d) {
    CNB_PDL_SYNC();
    __shared__ float red[TINY_WG_MAX_N * TINY_WG_MAX_C];
    const int tap = blockIdx.y;
    const int cbase = blockIdx.z * TINY_WG_MAX_C;
    const int cn = d.src_c - cbase < TINY_WG_MAX_C ? d.src_c - cbase : TINY_WG_MAX_C;  // channels of this chunk
    const int ky = tap / d.KW, kx = tap - ky * d.KW;
    if (threadIdx.x < TINY_WG_MAX_N * TINY_WG_MAX_C) red[threadIdx.x] = 0.f;
    __syncthreads();
    float acc[TINY_WG_MAX_N][TINY_WG_MAX_C];
#pragma unroll
    for (int n = 0; n < TINY_WG_MAX_N; ++n)
#pragma unroll
        for (int c = 0; c < TINY_WG_MAX_C; ++c) acc[n][c] = 0.f;
    const long M = (long)d.B * d.Hout * d.Wout;
    const T* src = reinterpret_cast<const T*>(d.src);
    const T* dy = reinterpret_cast<const T*>(d.dy);
    for (long m = (long)blockIdx.x * blockDim.x + threadIdx.x; m < M; m += (long)gridDim.x * blockDim.x) {
        const int ox = (int)(m % d.Wout);
        const long t = m / d.Wout;
        const int oy = (int)(t % d.Hout);
        const int ob = (int)(t / d.Hout);
        int iy, ix;
        if (!conv_src_coord(oy, ky, d.stride, d.pad, d.dil, d.transposed, d.Hin, iy)) continue;
        if (!conv_src_coord(ox, kx, d.stride, d.pad, d.dil, d.transposed, d.Win, ix)) continue;
        const T* sp = src + (((long)ob * d.Hin + iy) * d.Win + ix) * d.src_stride + cbase;
        float xs[TINY_WG_MAX_C];
#pragma unroll
        for (int c = 0; c < TINY_WG_MAX_C; ++c) xs[c] = c < cn ? cnb_ld(sp + c) : 0.f;
#pragma unroll
        for (int n = 0; n < TINY_WG_MAX_N; ++n) {
            if (n < d.N) {
                const float g = cnb_ld(dy + m * d.dy_stride + n);
#pragma unroll
                for (int c = 0; c < TINY_WG_MAX_C; ++c) acc[n][c] = fmaf(g, xs[c], acc[n][c]);
            }
        }
    }
#pragma unroll
    for (int n = 0; n < TINY_WG_MAX_N; ++n)
#pragma unroll
        for (int c = 0; c < TINY_WG_MAX_C; ++c) {
            if (n < d.N && c < cn) {  // uniform across the block
                const float v = cnb_warp_sum(acc[n][c]);
                if ((threadIdx.x & 31) == 0) atomicAdd(&red[n * TINY_WG_MAX_C + c], v);
            }
        }
    __syncthreads();
    if (threadIdx.x < TINY_WG_MAX_N * TINY_WG_MAX_C) {
        const int n = threadIdx.x / TINY_WG_MAX_C, c = threadIdx.x % TINY_WG_MAX_C;
        if (n < d.N && c < cn) atomicAdd(d.dwp + ((long)tap * d.N + n) * d.Ctot + d.k_off + cbase + c, red[threadIdx.x]);
    }
}

// The Psi-Net head shapes again ((N, C) = (3, 9) and (3, 3)): the kernel above walks dY and X once per (tap, 4-channel chunk) = 27
// passes over a 524 288-pixel level for the 9 -> 3 convolution (ncu: 85 us per launch for 13 MB of operands).  Here a thread keeps the
// N x C products of TG taps in registers (N*C*TG <= 81), so the level is walked taps / TG times (3 for 9 -> 3, once for 3 -> 3), dY
// is read once per pixel and pass, and the block reduces its TG*N*C partial sums once at the end.
template <typename T, int N, int C, int TG>
__global__ void __launch_bounds__(256) conv_tiny_wgrad_fixed_kernel(cnb_wgrad_desc d) {
    CNB_PDL_SYNC();
    __shared__ float red[TG * N * C];
    const int taps = d.KH * d.KW;
    const int tap0 = blockIdx.y * TG;
    for (int i = threadIdx.x; i < TG * N * C; i += blockDim.x) red[i] = 0.f;
    __syncthreads();
    float acc[TG][N][C];
#pragma unroll
    for (int j = 0; j < TG; ++j)
#pragma unroll
        for (int n = 0; n < N; ++n)
#pragma unroll
            for (int c = 0; c < C; ++c) acc[j][n][c] = 0.f;
    const long M = (long)d.B * d.Hout * d.Wout;
    const T* src = reinterpret_cast<const T*>(d.src);
    const T* dy = reinterpret_cast<const T*>(d.dy);
    for (long m = (long)blockIdx.x * blockDim.x + threadIdx.x; m < M; m += (long)gridDim.x * blockDim.x) {
        const int ox = (int)(m % d.Wout);
        const long t = m / d.Wout;
        const int oy = (int)(t % d.Hout);
        const int ob = (int)(t / d.Hout);
        float g[N];
#pragma unroll
        for (int n = 0; n < N; ++n) g[n] = cnb_ld(dy + m * d.dy_stride + n);
#pragma unroll
        for (int j = 0; j < TG; ++j) {
            const int tap = tap0 + j;
            if (tap >= taps) continue;
            const int ky = tap / d.KW, kx = tap - ky * d.KW;
            int iy, ix;
            if (!conv_src_coord(oy, ky, d.stride, d.pad, d.dil, d.transposed, d.Hin, iy)) continue;
            if (!conv_src_coord(ox, kx, d.stride, d.pad, d.dil, d.transposed, d.Win, ix)) continue;
            const T* sp = src + (((long)ob * d.Hin + iy) * d.Win + ix) * d.src_stride;
            float xs[C];
#pragma unroll
            for (int c = 0; c < C; ++c) xs[c] = cnb_ld(sp + c);
#pragma unroll
            for (int n = 0; n < N; ++n)
#pragma unroll
                for (int c = 0; c < C; ++c) acc[j][n][c] = fmaf(g[n], xs[c], acc[j][n][c]);
        }
    }
#pragma unroll
    for (int j = 0; j < TG; ++j)
#pragma unroll
        for (int n = 0; n < N; ++n)
#pragma unroll
            for (int c = 0; c < C; ++c) {
                const float v = cnb_warp_sum(acc[j][n][c]);
                if ((threadIdx.x & 31) == 0) atomicAdd(&red[(j * N + n) * C + c], v);
            }
    __syncthreads();
    for (int i = threadIdx.x; i < TG * N * C; i += blockDim.x) {
        const int c = i % C, n = (i / C) % N, tap = tap0 + i / (C * N);
        if (tap < taps) atomicAdd(d.dwp + ((long)tap * N + n) * d.Ctot + d.k_off + c, red[i]);
    }
}

// dst[p][0..C) = src[p][0..C), dst[p][C..dst_stride) = 0: gives a skinny tensor the 16-byte pixel pitch TMA needs
template <typename T>
__global__ void repitch_kernel(const T* __restrict__ src, int src_stride, T* __restrict__ dst, int dst_stride, long P, int C) {
    CNB_PDL_SYNC();
    const long total = P * dst_stride;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        const int c = (int)(i % dst_stride);
        const long p = i / dst_stride;
        dst[i] = c < C ? src[p * src_stride + c] : T(0.f);
    }
}

// Multi-tensor copy / accumulate of small fp32 segments: one launch for what would otherwise be one tiny torch kernel per segment
// (stacking the Psi-Net stream parameters, scattering BatchNorm / LayerNorm / scalar parameter gradients into the flat gradient
// buffer, writing stacked running statistics back).  grid = (blocks per segment, segments).
struct MultiCopyEntry {
    const float* src;
    float* dst;
    int n;
    int mode;  // 0: dst = src, 1: dst += src
};
// the segment table travels BY VALUE in the kernel arguments (160 x 24 bytes, inside the 4 KB argument space): nothing to upload, and a
// captured CUDA graph keeps its own copy
constexpr int MULTI_COPY_BATCH = 160;
struct MultiCopyBatch {
    MultiCopyEntry e[MULTI_COPY_BATCH];
};
__global__ void multi_copy_kernel(const __grid_constant__ MultiCopyBatch batch) {
    CNB_PDL_SYNC();
    const MultiCopyEntry& e = batch.e[blockIdx.y];
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < e.n; i += gridDim.x * blockDim.x) e.dst[i] = e.mode ? e.dst[i] + e.src[i] : e.src[i];
}

// out[b][p][c] = e[b][c]: a per-sample vector over the pixels of a level (GeoEmbeddings, reference unet_parts.py:742-750)
template <typename T>
__global__ void broadcast_pixels_kernel(const T* __restrict__ e, T* __restrict__ out, int B, long HW, int C) {
    CNB_PDL_SYNC();
    const long total = (long)B * HW * C;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        const int c = (int)(i % C);
        const long b = i / ((long)HW * C);
        out[i] = e[b * C + c];
    }
}

}  // namespace cnb
