// The two steps either side of predict_step on a resident satellite tile, both pure HBM streams:
//
//  * window_load_kernel  -- cuts prediction windows (window + halo) out of the int16 time-series tile and applies the load-time
//    arithmetic of the reference's dataset in the same pass: data/create.py:201-214 (rechunk to window_size, map_overlap with
//    depth = padding and boundary = 0), data/store.py:68-90 (ragged end chunks zero-padded after the data), data/datasets.py:443
//    (x / 10000 clipped to [1e-9, 1]) and utils/normalize.py:78-80 ((x - mean_c) / std_c).
//  * predict_pack_kernel -- what LightningGTiffWriter.write_on_batch_end does per window (callbacks.py:176-227): drop the halo,
//    stack (distance, edge, crop), x 10000, clip to [0, 10000], store as uint16 into the window's place of the 3-band mosaic.
//
// Divisions and products use the IEEE intrinsics (the library is built with --use_fast_math): results are bit-identical to the
// reference's fp32 torch / numpy arithmetic.
#pragma once
#include "cnb_common.cuh"

namespace cnb {

// Correctly rounded a / b from y = RN(1 / b) in five instructions instead of the ~12 + range check of div.rn.f32: q0 = RN(a y) is
// within 1.5 ulp of a / b, one residual step makes it faithful, and by Markstein's theorem a second residual step from a faithful
// quotient with a correctly rounded reciprocal gives RN(a / b) -- for divisors whose significand is not all ones and operands far
// from the overflow / underflow thresholds, which `markstein_ok` checks once per CTA (otherwise __fdiv_rn).
__device__ __forceinline__ bool markstein_ok(float b) {
    const float ab = fabsf(b);
    return ab > 1e-18f && ab < 1e18f && (__float_as_uint(b) & 0x7fffffu) != 0x7fffffu;
}
__device__ __forceinline__ float div_rn_markstein(float a, float b, float y) {
    float q = __fmul_rn(a, y);
    float r = __fmaf_rn(-q, b, a);
    q = __fmaf_rn(r, y, q);
    r = __fmaf_rn(-q, b, a);
    return __fmaf_rn(r, y, q);
}

// tile  int16 [T][C][Ht][Wt]  (reference order: time, band, y, x -- data/create.py:70-79)
// win   int32 [B][win_stride]  row b starts with (row_off, col_off), the window's un-padded origin in the tile
// out   fp32  [B][C][T][Hw][Ww], Hw = Ww = window_size + 2 * pad; (y, x) of the window is tile pixel (row_off - pad + y, col_off - pad + x)
// One CTA = one (b, c, t) plane of a window (grid = B*C*T, a multiple of waves at cfg 5: 1920 planes); one thread = four consecutive
// x of one row (Ww % 4 == 0): a 16-byte store, four 2-byte loads that share sectors with the neighbours.  The (row, quad) position
// advances incrementally -- no division inside the loop; band constants and the window origin are per-CTA values.
// VEC: the four int16 of a quad are one aligned 8-byte load (tile width, window column origin and base pointer multiples of 4
// elements -- true for 100 px windows with a 20 px halo on a 10980 px tile); otherwise four 2-byte loads.  Each thread first issues
// the loads of WL_UNROLL quads, then does the arithmetic: a 2-byte-per-element stream needs that many bytes in flight per thread
// to cover the HBM latency (the first version, one quad at a time, stopped at 3.2 TB/s).
#define WL_UNROLL 4
template <bool VEC>
__device__ __forceinline__ void window_plane(const int16_t* __restrict__ src_plane, float* __restrict__ dst_plane, int Ht, int Wt, int Hw,
                                             int Ww, int row0, int col00, float scale, float lo, float hi, bool mean, float m, bool stdv,
                                             float s) {
    const int quads = Ww >> 2;
    const int step_y = (int)blockDim.x / quads, step_x = (int)blockDim.x % quads;
    int y = (int)threadIdx.x / quads, xq = (int)threadIdx.x % quads;
    // |raw| <= 32768 and the clipped reflectances are far inside the normal range: only the divisors decide
    const bool fast_div = markstein_ok(scale) && markstein_ok(s) && fabsf(m) < 1e18f && fabsf(lo) < 1e18f && fabsf(hi) < 1e18f;
    const float y_scale = __frcp_rn(scale), y_s = __frcp_rn(s);
    while (y < Hw) {
        int ys[WL_UNROLL], xs[WL_UNROLL];
        short4 raw[WL_UNROLL];
#pragma unroll
        for (int u = 0; u < WL_UNROLL; ++u) {
            ys[u] = y, xs[u] = xq;
            y += step_y, xq += step_x;
            if (xq >= quads) xq -= quads, ++y;
        }
#pragma unroll
        for (int u = 0; u < WL_UNROLL; ++u) {
            raw[u] = make_short4(0, 0, 0, 0);
            const int row = row0 + ys[u], col0 = col00 + 4 * xs[u];
            if (ys[u] < Hw && row >= 0 && row < Ht) {
                const int16_t* src = src_plane + (long)row * Wt + col0;
                if (VEC) {
                    if (col0 >= 0 && col0 < Wt) raw[u] = *reinterpret_cast<const short4*>(src);  // quads are entirely in or out
                } else {
                    if (col0 >= 0 && col0 < Wt) raw[u].x = src[0];
                    if (col0 + 1 >= 0 && col0 + 1 < Wt) raw[u].y = src[1];
                    if (col0 + 2 >= 0 && col0 + 2 < Wt) raw[u].z = src[2];
                    if (col0 + 3 >= 0 && col0 + 3 < Wt) raw[u].w = src[3];
                }
            }
        }
#pragma unroll
        for (int u = 0; u < WL_UNROLL; ++u) {
            if (ys[u] >= Hw) break;
            float v[4] = {(float)raw[u].x, (float)raw[u].y, (float)raw[u].z, (float)raw[u].w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                float q = fast_div ? div_rn_markstein(v[j], scale, y_scale) : __fdiv_rn(v[j], scale);
                q = fminf(fmaxf(q, lo), hi);
                if (mean) q = __fsub_rn(q, m);
                if (stdv) q = fast_div ? div_rn_markstein(q, s, y_s) : __fdiv_rn(q, s);
                v[j] = q;
            }
            *reinterpret_cast<float4*>(dst_plane + (long)ys[u] * Ww + 4 * xs[u]) = make_float4(v[0], v[1], v[2], v[3]);
        }
    }
}

__global__ void __launch_bounds__(256) window_load_kernel(const int16_t* __restrict__ tile, int T, int C, int Ht, int Wt,
                                                          const int32_t* __restrict__ win, int win_stride, int B, int Hw, int Ww, int pad,
                                                          float scale, float lo, float hi, const float* __restrict__ mean, const float* __restrict__ stdv,
                                                          float* __restrict__ out, int vec_ok) {
    CNB_PDL_SYNC();
    const int plane = blockIdx.x;  // ((b * C) + c) * T + t
    const int t = plane % T;
    const int c = (plane / T) % C;
    const int b = plane / (T * C);
    const int row0 = win[win_stride * b] - pad, col00 = win[win_stride * b + 1] - pad;
    const float m = mean ? mean[c] : 0.f;
    const float s = stdv ? stdv[c] : 1.f;
    const int16_t* src_plane = tile + ((long)t * C + c) * Ht * Wt;
    float* dst_plane = out + (long)plane * Hw * Ww;
    if (vec_ok && (col00 & 3) == 0)  // uniform over the CTA
        window_plane<true>(src_plane, dst_plane, Ht, Wt, Hw, Ww, row0, col00, scale, lo, hi, mean != nullptr, m, stdv != nullptr, s);
    else
        window_plane<false>(src_plane, dst_plane, Ht, Wt, Hw, Ww, row0, col00, scale, lo, hi, mean != nullptr, m, stdv != nullptr, s);
}

// dist / edge / crop: fp32, element (b, y, x) at ptr[b * batch_stride + y * Ws + x] (the [B,1,Hs,Ws] outputs of predict_step; a
//                     multi-class crop output passes channel 1 through the pointer offset, callbacks.py:131-132)
// win   int32 [B][4]  (row_off, col_off, height, width); height/width are clipped to the mosaic here as callbacks.py:182-185 does;
//                     height = 0 marks a filler window of a ragged last batch
// mosaic uint16 [3][Ht][pitch >= Wt]
// One CTA = one output row of one band of one window (grid = B * 3 * win_size).
__global__ void __launch_bounds__(128) predict_pack_kernel(const float* __restrict__ dist, const float* __restrict__ edge,
                                                           const float* __restrict__ crop, long batch_stride, int Hs, int Ws, int pad,
                                                           const int32_t* __restrict__ win, int B, int win_size, float scale,
                                                           uint16_t* __restrict__ mosaic, int Ht, int Wt, int pitch) {
    CNB_PDL_SYNC();
    const int y = blockIdx.x % win_size;
    const int pl = blockIdx.x / win_size;
    const int band = pl % 3, b = pl / 3;
    const int row_off = win[4 * b], col_off = win[4 * b + 1];
    int h = win[4 * b + 2], w = win[4 * b + 3];
    if (row_off + h > Ht) h = Ht - row_off;
    if (col_off + w > Wt) w = Wt - col_off;
    if (y >= h || pad + y >= Hs) return;
    if (w > Ws - pad) w = Ws - pad;
    const float* src = (band == 0 ? dist : (band == 1 ? edge : crop)) + b * batch_stride + (long)(pad + y) * Ws + pad;
    uint16_t* dst = mosaic + ((long)band * Ht + row_off + y) * pitch + col_off;
    for (int x = threadIdx.x; x < w; x += blockDim.x) {
        float v = __fmul_rn(src[x], scale);
        v = fminf(fmaxf(v, 0.f), scale);  // NaN -> 0 (fmaxf returns the non-NaN operand)
        dst[x] = (uint16_t)(int)v;        // C truncation = numpy's astype
    }
}

}  // namespace cnb
