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

// tile  int16 [T][C][Ht][Wt]  (reference order: time, band, y, x -- data/create.py:70-79)
// win   int32 [B][win_stride]  row b starts with (row_off, col_off), the window's un-padded origin in the tile
// out   fp32  [B][C][T][Hw][Ww], Hw = Ww = window_size + 2 * pad; (y, x) of the window is tile pixel (row_off - pad + y, col_off - pad + x)
// One thread = four consecutive x of one row (Ww % 4 == 0): a 16-byte store, four 2-byte loads that share sectors with the neighbours.
__global__ void __launch_bounds__(256) window_load_kernel(const int16_t* __restrict__ tile, int T, int C, int Ht, int Wt,
                                                          const int32_t* __restrict__ win, int win_stride, int B, int Hw, int Ww, int pad,
                                                          float scale, float lo, float hi, const float* __restrict__ mean, const float* __restrict__ stdv,
                                                          float* __restrict__ out) {
    CNB_PDL_SYNC();
    const int quads = Ww >> 2;
    const long total = (long)B * C * T * Hw * quads;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        const int xq = (int)(i % quads);
        long r = i / quads;
        const int y = (int)(r % Hw);
        r /= Hw;
        const int t = (int)(r % T);
        r /= T;
        const int c = (int)(r % C);
        const int b = (int)(r / C);
        const int row = win[win_stride * b] - pad + y;
        const int col0 = win[win_stride * b + 1] - pad + 4 * xq;
        const float m = mean ? mean[c] : 0.f;
        const float s = stdv ? stdv[c] : 1.f;
        const bool row_ok = row >= 0 && row < Ht;
        const int16_t* src = tile + (((long)t * C + c) * Ht + (row_ok ? row : 0)) * Wt;
        float v[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int col = col0 + j;
            const float raw = (row_ok && col >= 0 && col < Wt) ? (float)src[col] : 0.f;
            float q = __fdiv_rn(raw, scale);
            q = fminf(fmaxf(q, lo), hi);
            if (mean) q = __fsub_rn(q, m);
            if (stdv) q = __fdiv_rn(q, s);
            v[j] = q;
        }
        *reinterpret_cast<float4*>(out + i * 4) = make_float4(v[0], v[1], v[2], v[3]);
    }
}

// dist / edge / crop: fp32, element (b, y, x) at ptr[b * batch_stride + y * Ws + x] (the [B,1,Hs,Ws] outputs of predict_step; a
//                     multi-class crop output passes channel 1 through the pointer offset, callbacks.py:131-132)
// win   int32 [B][4]  (row_off, col_off, height, width); height/width are clipped to the mosaic here as callbacks.py:182-185 does;
//                     height = 0 marks a filler window of a ragged last batch
// mosaic uint16 [3][Ht][pitch >= Wt]
__global__ void __launch_bounds__(256) predict_pack_kernel(const float* __restrict__ dist, const float* __restrict__ edge,
                                                           const float* __restrict__ crop, long batch_stride, int Hs, int Ws, int pad,
                                                           const int32_t* __restrict__ win, int B, int win_size, float scale,
                                                           uint16_t* __restrict__ mosaic, int Ht, int Wt, int pitch) {
    CNB_PDL_SYNC();
    const long per_b = 3L * win_size * win_size;
    const long total = (long)B * per_b;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        const int x = (int)(i % win_size);
        long r = i / win_size;
        const int y = (int)(r % win_size);
        r /= win_size;
        const int band = (int)(r % 3);
        const int b = (int)(r / 3);
        const int row_off = win[4 * b], col_off = win[4 * b + 1];
        int h = win[4 * b + 2], w = win[4 * b + 3];
        if (row_off + h > Ht) h = Ht - row_off;
        if (col_off + w > Wt) w = Wt - col_off;
        if (y >= h || x >= w || pad + y >= Hs || pad + x >= Ws) continue;
        const float* src = band == 0 ? dist : (band == 1 ? edge : crop);
        float v = __fmul_rn(src[b * batch_stride + (long)(pad + y) * Ws + (pad + x)], scale);
        v = fminf(fmaxf(v, 0.f), scale);  // NaN -> 0 (fmaxf returns the non-NaN operand)
        mosaic[((long)band * Ht + row_off + y) * pitch + col_off + x] = (uint16_t)(int)v;  // C truncation = numpy's astype
    }
}

}  // namespace cnb
