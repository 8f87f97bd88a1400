// Generic implicit-GEMM convolution (CUDA cores, fp32 accumulate) over pixel-major sources.
//
// This is the correctness baseline and the path for shapes the tcgen05 kernel does not take
// (channel counts that are not multiples of 64, strided / transposed gathers, fp32 parity mode).
// One CTA computes a 64-pixel x 64-channel output tile; K runs over (tap, source, 16-channel chunk).
#pragma once
#include "cnb_common.cuh"

namespace cnb {

struct ConvGeom {
    int B, Hin, Win, Hout, Wout;
    int KH, KW, stride, pad, dil, transposed;
};

// source coordinate of output coordinate `o` for kernel tap `k`; returns false when the tap falls outside
__device__ __forceinline__ bool conv_src_coord(int o, int k, int stride, int pad, int dil, int transposed, int in_len, int& i) {
    if (!transposed) {
        i = o * stride - pad + k * dil;
        return i >= 0 && i < in_len;
    }
    int num = o + pad - k * dil;
    if (num < 0) return false;
    i = num / stride;
    return (num - i * stride) == 0 && i < in_len;
}

constexpr int CG_BM = 64, CG_BN = 64, CG_BK = 16, CG_PAD = 4;

template <typename T>
__global__ void __launch_bounds__(256) conv_fwd_generic_kernel(cnb_conv_desc d) {
    CNB_PDL_SYNC();
    __shared__ float As[CG_BK][CG_BM + CG_PAD];
    __shared__ float Bs[CG_BK][CG_BN + CG_PAD];
    const int tid = threadIdx.x;
    const long M = (long)d.B * d.Hout * d.Wout;
    const long m0 = (long)blockIdx.x * CG_BM;
    const int n0 = blockIdx.y * CG_BN;

    // loader mapping: 64 rows x 4 quads of 4 consecutive k
    const int l_row = tid >> 2;
    const int l_k = (tid & 3) * 4;
    const long m = m0 + l_row;
    const bool m_ok = m < M;
    int ob = 0, oy = 0, ox = 0;
    if (m_ok) {
        ox = (int)(m % d.Wout);
        long t = m / d.Wout;
        oy = (int)(t % d.Hout);
        ob = (int)(t / d.Hout);
    }
    const int bn = n0 + l_row;
    const bool n_ok = bn < d.N;

    const int ty = tid >> 4, tx = tid & 15;
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    const T* wbase = reinterpret_cast<const T*>(d.w_packed);
    const int taps = d.KH * d.KW;
    for (int tap = 0; tap < taps; ++tap) {
        const int ky = tap / d.KW, kx = tap - ky * d.KW;
        int iy = 0, ix = 0;
        bool valid = m_ok;
        if (valid) valid = conv_src_coord(oy, ky, d.stride, d.pad, d.dil, d.transposed, d.Hin, iy);
        if (valid) valid = conv_src_coord(ox, kx, d.stride, d.pad, d.dil, d.transposed, d.Win, ix);
        const long pix = valid ? (((long)ob * d.Hin + iy) * d.Win + ix) : 0;
        int coff = 0;
        for (int s = 0; s < d.nsrc; ++s) {
            const int Cs = d.src_c[s];
            const T* sp = reinterpret_cast<const T*>(d.src[s]) + pix * d.src_stride[s];
            const T* wp = wbase + (long)tap * d.w_tap_stride + (long)bn * d.w_row_stride + coff;
            for (int c0 = 0; c0 < Cs; c0 += CG_BK) {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int c = c0 + l_k + j;
                    float av = 0.f, bv = 0.f;
                    if (valid && c < Cs) av = cnb_ld(sp + c);
                    if (n_ok && c < Cs) bv = cnb_ld(wp + c);
                    As[l_k + j][l_row] = av;
                    Bs[l_k + j][l_row] = bv;
                }
                __syncthreads();
#pragma unroll
                for (int kk = 0; kk < CG_BK; ++kk) {
                    float a[4], b[4];
#pragma unroll
                    for (int i = 0; i < 4; ++i) a[i] = As[kk][ty * 4 + i];
#pragma unroll
                    for (int j = 0; j < 4; ++j) b[j] = Bs[kk][tx * 4 + j];
#pragma unroll
                    for (int i = 0; i < 4; ++i)
#pragma unroll
                        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
                }
                __syncthreads();
            }
            coff += Cs;
        }
    }

    T* out = reinterpret_cast<T*>(d.out);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const long mm = m0 + ty * 4 + i;
        if (mm >= M) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int n = n0 + tx * 4 + j;
            if (n >= d.N) continue;
            float v = acc[i][j];
            if (d.bias) v += d.bias[n];
            cnb_st(out + mm * d.out_stride + n, v);
        }
    }
}

// dWp[tap][n][k_off + c] += sum over a slice of the pixels.  grid = (ceil(Cs/64), ceil(N/64), taps*splits)
template <typename T>
__global__ void __launch_bounds__(256) conv_wgrad_generic_kernel(cnb_wgrad_desc d, int splits, long m_per_split) {
    CNB_PDL_SYNC();
    __shared__ float As[CG_BK][CG_BM + CG_PAD];  // [pixel in chunk][source channel]
    __shared__ float Bs[CG_BK][CG_BN + CG_PAD];  // [pixel in chunk][output channel]
    const int tid = threadIdx.x;
    const long M = (long)d.B * d.Hout * d.Wout;
    const int c0 = blockIdx.x * CG_BM;
    const int n0 = blockIdx.y * CG_BN;
    const int tap = blockIdx.z / splits;
    const int split = blockIdx.z - tap * splits;
    const long m_begin = (long)split * m_per_split;
    long m_end = m_begin + m_per_split;
    if (m_end > M) m_end = M;
    const int ky = tap / d.KW, kx = tap - ky * d.KW;

    // loader mapping: 16 pixels x 16 quads of 4 consecutive channels
    const int l_p = tid >> 4;
    const int l_c = (tid & 15) * 4;
    const int ty = tid >> 4, tx = tid & 15;
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    const T* src = reinterpret_cast<const T*>(d.src);
    const T* dy = reinterpret_cast<const T*>(d.dy);
    for (long mc = m_begin; mc < m_end; mc += CG_BK) {
        const long m = mc + l_p;
        bool valid = m < m_end;
        int iy = 0, ix = 0, ob = 0;
        if (valid) {
            const int ox = (int)(m % d.Wout);
            const long t = m / d.Wout;
            const int oy = (int)(t % d.Hout);
            ob = (int)(t / d.Hout);
            valid = conv_src_coord(oy, ky, d.stride, d.pad, d.dil, d.transposed, d.Hin, iy) &&
                    conv_src_coord(ox, kx, d.stride, d.pad, d.dil, d.transposed, d.Win, ix);
        }
        const long pix = valid ? (((long)ob * d.Hin + iy) * d.Win + ix) : 0;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int c = c0 + l_c + j;
            const int n = n0 + l_c + j;
            float av = 0.f, bv = 0.f;
            if (valid && c < d.src_c) av = cnb_ld(src + pix * d.src_stride + c);
            if (valid && n < d.N) bv = cnb_ld(dy + m * d.dy_stride + n);
            As[l_p][l_c + j] = av;
            Bs[l_p][l_c + j] = bv;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < CG_BK; ++kk) {
            float a[4], b[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) a[i] = As[kk][ty * 4 + i];
#pragma unroll
            for (int j = 0; j < 4; ++j) b[j] = Bs[kk][tx * 4 + j];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
    }

    // acc[i][j]: source channel c0+ty*4+i, output channel n0+tx*4+j
    float* dw = d.dwp + (long)tap * d.N * d.Ctot;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int n = n0 + tx * 4 + j;
        if (n >= d.N) continue;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int c = c0 + ty * 4 + i;
            if (c >= d.src_c) continue;
            atomicAdd(dw + (long)n * d.Ctot + d.k_off + c, acc[i][j]);
        }
    }
}

template <typename T>
__global__ void pack_weight_kernel(const float* __restrict__ w, T* __restrict__ wp, int taps, int N, int K, int pitch, long s_n, long s_k,
                                   long s_tap) {
    CNB_PDL_SYNC();
    const long total = (long)taps * N * pitch;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        const int k = (int)(i % pitch);
        const long t = i / pitch;
        const int n = (int)(t % N);
        const int tap = (int)(t / N);
        cnb_st(wp + i, k < K ? w[n * s_n + k * s_k + tap * s_tap] : 0.f);
    }
}

__global__ void unpack_wgrad_kernel(float* __restrict__ dwp, float* __restrict__ g, int taps, int N, int K, long s_n, long s_k,
                                    long s_tap, int mode) {
    CNB_PDL_SYNC();
    const long total = (long)taps * N * K;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        const int k = (int)(i % K);
        const long t = i / K;
        const int n = (int)(t % N);
        const int tap = (int)(t / N);
        const long o = n * s_n + k * s_k + tap * s_tap;
        const float v = dwp[i];
        if (mode & 2) dwp[i] = 0.f;
        g[o] = (mode & 1) ? g[o] + v : v;
    }
}

// Tiled forms of the two kernels above.  A CTA owns a 32 (n) x 32 (k) tile of the parameter for all taps in shared memory, reads
// the fp32 parameter in ITS memory order (the innermost pair of (n, k) and the taps are contiguous floats) and writes BOTH packed
// layouts -- forward [tap][N][k-pitch] and data-gradient [tap][K][n-pitch] -- with the fastest index across the lanes, so the
// parameter is read once per step instead of twice and no access is a 36-byte-strided scalar (the flat kernels: 8x sector
// over-fetch on the gather, one launch per layout).  Padding columns (k >= K resp. n >= N up to the pitch) are written as zero.
constexpr int PW_T = 32;
template <typename T>
__device__ __forceinline__ void pack_weight_tile(float* tile, const float* __restrict__ w, T* __restrict__ wp, T* __restrict__ wd, int taps,
                                                 int N, int K, int pitch_k, int pitch_n, long s_n, long s_k, long s_tap, int n0, int k0) {
    const int per = PW_T * taps;  // contiguous floats per outer index when s_tap == 1 and the inner stride == taps
    const bool k_inner = s_k <= s_n;
#pragma unroll 4
    for (int i = threadIdx.x; i < PW_T * per; i += blockDim.x) {
        const int outer = i / per, rem = i - outer * per;
        const int inner = rem / taps, tap = rem - inner * taps;
        const int nl = k_inner ? outer : inner, kl = k_inner ? inner : outer;
        const int n = n0 + nl, k = k0 + kl;
        tile[(tap * PW_T + nl) * (PW_T + 1) + kl] = (n < N && k < K) ? w[n * s_n + k * s_k + tap * s_tap] : 0.f;
    }
    __syncthreads();
#pragma unroll 4
    for (int i = threadIdx.x; i < taps * PW_T * PW_T; i += blockDim.x) {
        const int kl = i % PW_T, t = i / PW_T;
        const int nl = t % PW_T, tap = t / PW_T;
        const int n = n0 + nl, k = k0 + kl;
        if (n < N && k < pitch_k) cnb_st(wp + ((long)tap * N + n) * pitch_k + k, tile[(tap * PW_T + nl) * (PW_T + 1) + kl]);
    }
    if (wd) {
#pragma unroll 4
        for (int i = threadIdx.x; i < taps * PW_T * PW_T; i += blockDim.x) {
            const int nl = i % PW_T, t = i / PW_T;
            const int kl = t % PW_T, tap = t / PW_T;
            const int n = n0 + nl, k = k0 + kl;
            if (k < K && n < pitch_n) cnb_st(wd + ((long)tap * K + k) * pitch_n + n, tile[(tap * PW_T + nl) * (PW_T + 1) + kl]);
        }
    }
}

template <typename T>
__global__ void __launch_bounds__(256) pack_weight_tiled_kernel(const float* __restrict__ w, T* __restrict__ wp, T* __restrict__ wd, int taps,
                                                               int N, int K, int pitch_k, int pitch_n, long s_n, long s_k, long s_tap) {
    CNB_PDL_SYNC();
    CNB_DYN_SMEM(sm_raw);  // float tile[taps][PW_T][PW_T + 1]
    pack_weight_tile<T>(reinterpret_cast<float*>(sm_raw), w, wp, wd, taps, N, K, pitch_k, pitch_n, s_n, s_k, s_tap, blockIdx.y * PW_T,
                        blockIdx.x * PW_T);
}

// Every convolution weight of the model in ONE launch: a device-resident table of cnb_pack_desc (sorted by first tile index) and
// one CTA per 32 x 32 tile of any of them.  96 separate launches of 64-320 CTAs each cost 2.1 ms per step (ncu), the work is ~0.1 ms.
template <typename T>
__global__ void __launch_bounds__(256) pack_weight_batched_kernel(const cnb_pack_desc* __restrict__ table, int ndesc) {
    CNB_PDL_SYNC();
    CNB_DYN_SMEM(sm_raw);
    int lo = 0, hi = ndesc - 1;  // last descriptor whose tile0 <= blockIdx.x
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (table[mid].tile0 <= (int)blockIdx.x)
            lo = mid;
        else
            hi = mid - 1;
    }
    const cnb_pack_desc d = table[lo];
    const int t = (int)blockIdx.x - d.tile0;
    pack_weight_tile<T>(reinterpret_cast<float*>(sm_raw), d.w, reinterpret_cast<T*>(d.wp), reinterpret_cast<T*>(d.wd), d.taps, d.N, d.K,
                        d.pitch_k, d.pitch_n, (long)d.s_n, (long)d.s_k, (long)d.s_tap, (t / d.tiles_x) * PW_T, (t % d.tiles_x) * PW_T);
}

// g[n*s_n + k*s_k + tap*s_tap] (+)= dwp[tap][n][k], written in the parameter's memory order.  mode bit 0: accumulate into g; bit 1:
// clear dwp after reading it (a persistent per-parameter accumulator is then zero again for the next step: no fill kernel).
__device__ __forceinline__ void unpack_wgrad_tile(float* tile, float* __restrict__ dwp, float* __restrict__ g, int taps, int N, int K, long s_n,
                                                  long s_k, long s_tap, int mode, int n0, int k0) {
    const bool accumulate = mode & 1, clear = mode & 2;
    // read phase: thread (kl = t % 32, nl0 = t / 32) owns column k of rows nl0, nl0 + 8, .. of every tap: no divisions; the four row
    // loads of a tap are issued together, and the accumulator is cleared only after them (a store between the loads would serialise
    // them on the load latency: the compiler cannot prove that the stores do not alias the next loads)
    {
        const int kl = threadIdx.x % PW_T, nl0 = threadIdx.x / PW_T;  // PW_T = 32, 256 threads: 8 rows per pass
        const int k = k0 + kl;
        constexpr int ROWS = PW_T / (256 / PW_T);  // 4
#pragma unroll 2
        for (int tap = 0; tap < taps; ++tap) {
            float v[ROWS];
#pragma unroll
            for (int r = 0; r < ROWS; ++r) {
                const int n = n0 + nl0 + r * (256 / PW_T);
                v[r] = (n < N && k < K) ? dwp[((long)tap * N + n) * K + k] : 0.f;
            }
#pragma unroll
            for (int r = 0; r < ROWS; ++r) tile[(tap * PW_T + nl0 + r * (256 / PW_T)) * (PW_T + 1) + kl] = v[r];
        }
        if (clear)
            for (int tap = 0; tap < taps; ++tap)
#pragma unroll
                for (int r = 0; r < ROWS; ++r) {
                    const int n = n0 + nl0 + r * (256 / PW_T);
                    if (n < N && k < K) dwp[((long)tap * N + n) * K + k] = 0.f;
                }
    }
    __syncthreads();
    const int per = PW_T * taps;
    const bool k_inner = s_k <= s_n;
    // write phase: position p = (inner, tap) inside a run of `per` consecutive outputs; p advances by 256 per step, (inner, tap) follow
    // incrementally
    const int step_inner = 256 / taps, step_tap = 256 % taps;
    for (int outer = 0; outer < PW_T; ++outer) {
        int inner = threadIdx.x / taps, tap = threadIdx.x % taps;
        for (int pp = threadIdx.x; pp < per; pp += 256) {
            const int nl = k_inner ? outer : inner, kl = k_inner ? inner : outer;
            const int n = n0 + nl, k = k0 + kl;
            if (n < N && k < K) {
                const long o = n * s_n + k * s_k + tap * s_tap;
                const float v = tile[(tap * PW_T + nl) * (PW_T + 1) + kl];
                g[o] = accumulate ? g[o] + v : v;
            }
            inner += step_inner;
            tap += step_tap;
            if (tap >= taps) tap -= taps, ++inner;
        }
    }
}

__global__ void __launch_bounds__(256) unpack_wgrad_tiled_kernel(float* __restrict__ dwp, float* __restrict__ g, int taps, int N, int K,
                                                                long s_n, long s_k, long s_tap, int mode) {
    CNB_PDL_SYNC();
    CNB_DYN_SMEM(sm_raw);
    unpack_wgrad_tile(reinterpret_cast<float*>(sm_raw), dwp, g, taps, N, K, s_n, s_k, s_tap, mode, blockIdx.y * PW_T, blockIdx.x * PW_T);
}

// Every weight gradient of a backward pass in ONE launch (the inverse of pack_weight_batched_kernel, same descriptor table layout:
// w = gradient destination, wp = fp32 accumulator [taps][N][K], reserved = mode).  103 separate launches of 4-320 CTAs cost 1.1 ms per
// step; the work is 0.5 GB of traffic.
__global__ void __launch_bounds__(256) unpack_wgrad_batched_kernel(const cnb_pack_desc* __restrict__ table, int ndesc) {
    CNB_PDL_SYNC();
    CNB_DYN_SMEM(sm_raw);
    int lo = 0, hi = ndesc - 1;  // last descriptor whose tile0 <= blockIdx.x
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (table[mid].tile0 <= (int)blockIdx.x)
            lo = mid;
        else
            hi = mid - 1;
    }
    const cnb_pack_desc d = table[lo];
    const int t = (int)blockIdx.x - d.tile0;
    unpack_wgrad_tile(reinterpret_cast<float*>(sm_raw), reinterpret_cast<float*>(d.wp), const_cast<float*>(d.w), d.taps, d.N, d.K, (long)d.s_n,
                      (long)d.s_k, (long)d.s_tap, d.reserved, (t / d.tiles_x) * PW_T, (t % d.tiles_x) * PW_T);
}

// db[n] += sum_p dy[p][n]; grid = (ceil(N/32), pixel splits); 8 pixel lanes per channel column, one atomic per CTA column
template <typename T>
__global__ void __launch_bounds__(256) bias_grad_kernel(const T* __restrict__ dy, int dy_stride, long P, int N, float* __restrict__ db) {
    CNB_PDL_SYNC();
    __shared__ float red[8][33];
    const int cx = threadIdx.x & 31, py = threadIdx.x >> 5;
    const int n = blockIdx.x * 32 + cx;
    float s = 0.f;
    if (n < N)
        for (long p = (long)blockIdx.y * 8 + py; p < P; p += (long)gridDim.y * 8) s += cnb_ld(dy + p * dy_stride + n);
    red[py][cx] = s;
    __syncthreads();
    if (py == 0 && n < N) {
        float t = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) t += red[i][cx];
        atomicAdd(db + n, t);
    }
}

}  // namespace cnb
