// Specialised neighbourhood-attention kernels for the shapes the TowerUNet runs in throughput mode (bf16, kernel 3 or 7,
// dilation 1 or 2, head_dim 32 or 64).  Same tiling as the generic tiled kernels of k_na.cuh (one CTA = an 8 x 16 pixel tile
// of one head, k/v or q/dout rows of tile + halo staged once in shared memory with cp.async) but:
//   * kernel size and dilation are template parameters: the k*k loops are fully unrolled, the logits stay in registers and the
//     softmax is evaluated once (one exp per neighbour instead of two plus a rescale of the accumulator), window starts
//     need no runtime division, every operand address is `base + compile-time multiple of two strides`;
//   * every k/v access is a shared-memory load (the staged region provably covers every window of the tile when
//     TILE >= dilation, see na2d_fwd_tile_kernel) and the lanes of a (pixel, head) group reduce with full-warp xor shuffles --
//     ncu/SASS of the generic kernel showed ~110 instructions per neighbour, half of them generic-address LD, 64-bit address
//     arithmetic and MATCH/VOTE sequences guarding partial-mask shuffles;
//   * the backward no longer recomputes q.k and dout.v in the key-side pass: the query-side pass stores p_in and
//     scale * p_in (dp_in - D_i) (72 bytes per (pixel, head) for k = 3), and the key-side pass is a pure gather
//         dk_j = sum_i ds_ij q_i,   dv_j = sum_i p_ij dout_i
//     with no dot products, no exponentials and no shuffles.
#pragma once
#include "cnb_common.cuh"
#include "k_na.cuh"

namespace cnb {
namespace naf {

template <int DIL>
__device__ __forceinline__ int wstart(int index, int length, int ksize, int dil_rt) {
    const int d = DIL > 0 ? DIL : dil_rt;
    const int g = index % d, p = index / d;
    const int group_len = (length - g + d - 1) / d;
    int s = p - ksize / 2;
    s = s < 0 ? 0 : s;
    s = s > group_len - ksize ? group_len - ksize : s;
    return g + d * s;
}

// sum over the LPH lanes of a (pixel, head) group; every lane of the warp takes part (uniform control flow by construction)
template <int LPH>
__device__ __forceinline__ float gsum(float v) {
#pragma unroll
    for (int o = LPH / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__device__ __forceinline__ void unpack8(const uint4& r, float* v) {
    v[0] = cnb_bits2f(r.x << 16), v[1] = cnb_bits2f(r.x & 0xffff0000u);
    v[2] = cnb_bits2f(r.y << 16), v[3] = cnb_bits2f(r.y & 0xffff0000u);
    v[4] = cnb_bits2f(r.z << 16), v[5] = cnb_bits2f(r.z & 0xffff0000u);
    v[6] = cnb_bits2f(r.w << 16), v[7] = cnb_bits2f(r.w & 0xffff0000u);
}
__device__ __forceinline__ float dot8(const float* a, const uint4& r) {
    float b[8];
    unpack8(r, b);
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) s = fmaf(a[j], b[j], s);
    return s;
}
__device__ __forceinline__ void axpy8(float w, const uint4& r, float* acc) {
    float b[8];
    unpack8(r, b);
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = fmaf(w, b[j], acc[j]);
}

struct TilePos {
    int head, b, y0, x0, ry0, rx0;
    long img_pix0;
};
__device__ __forceinline__ TilePos tile_pos(const NaTile& g) {
    TilePos t;
    t.head = blockIdx.y;
    int i = blockIdx.x;
    const int tx = i % g.tiles_x;
    i /= g.tiles_x;
    const int ty = i % g.tiles_y;
    t.b = i / g.tiles_y;
    const int halo = (g.ksize / 2) * g.dil;
    t.y0 = ty * NA_TH, t.x0 = tx * NA_TW;
    t.ry0 = na_region_origin(t.y0, halo, g.H, g.RH), t.rx0 = na_region_origin(t.x0, halo, g.W, g.RW);
    t.img_pix0 = (long)t.b * g.H * g.W;
    return t;
}

// ---------------------------------------------------------------------------------------------------------------------
// forward: out_i = softmax_n(scale q_i . k_n) v_n, lse_i
// ---------------------------------------------------------------------------------------------------------------------
template <int KS, int DIL, int LPH>
__global__ void __launch_bounds__(NA_TILE_THREADS) na2d_fwd_fast_kernel(const bf16_t* __restrict__ qkv, bf16_t* __restrict__ out,
                                                                       float* __restrict__ lse, NaTile g) {
    CNB_PDL_SYNC();
    constexpr int HD = LPH * 8, K2 = KS * KS;
    CNB_DYN_SMEM(sm_raw);
    bf16_t* sm = reinterpret_cast<bf16_t*>(sm_raw);
    const TilePos t = tile_pos(g);
    const int C = g.heads * HD;
    na_stage_region(sm, qkv + C + t.head * HD, qkv + 2 * C + t.head * HD, 3L * C, g, t.img_pix0, t.ry0, t.rx0);
    __syncthreads();
    const int dil = DIL > 0 ? DIL : g.dil;
    const int col_step = dil * 2 * HD, row_step = col_step * g.RW;  // elements between window columns / rows in the staged region

    for (int it = threadIdx.x; it < NA_TH * NA_TW * LPH; it += NA_TILE_THREADS) {
        const int sub = it % LPH, pl = it / LPH;
        const int lx = pl % NA_TW, ly = pl / NA_TW;
        const bool valid = (t.y0 + ly) < g.H && (t.x0 + lx) < g.W;
        // out-of-image lanes shadow the last pixel of the image (it lies in this tile): control flow and shuffles stay uniform
        const int y = (t.y0 + ly) < g.H ? t.y0 + ly : g.H - 1, x = (t.x0 + lx) < g.W ? t.x0 + lx : g.W - 1;
        const long pix = t.img_pix0 + (long)y * g.W + x;
        float q[8];
        unpack8(*reinterpret_cast<const uint4*>(qkv + pix * 3 * C + t.head * HD + sub * 8), q);
#pragma unroll
        for (int j = 0; j < 8; ++j) q[j] *= g.scale;
        const int sy = wstart<DIL>(y, g.H, KS, g.dil), sx = wstart<DIL>(x, g.W, KS, g.dil);
        const bf16_t* kb = sm + ((sy - t.ry0) * g.RW + (sx - t.rx0)) * 2 * HD + sub * 8;
        float lg[K2];
        float m = -INFINITY;
#pragma unroll
        for (int a = 0; a < KS; ++a)
#pragma unroll
            for (int b = 0; b < KS; ++b) {
                const float s = gsum<LPH>(dot8(q, *reinterpret_cast<const uint4*>(kb + a * row_step + b * col_step)));
                lg[a * KS + b] = s;
                m = fmaxf(m, s);
            }
        float l = 0.f;
#pragma unroll
        for (int n = 0; n < K2; ++n) {
            lg[n] = cnb_exp(lg[n] - m);
            l += lg[n];
        }
        float o[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = 0.f;
#pragma unroll
        for (int a = 0; a < KS; ++a)
#pragma unroll
            for (int b = 0; b < KS; ++b) axpy8(lg[a * KS + b], *reinterpret_cast<const uint4*>(kb + a * row_step + b * col_step + HD), o);
        if (valid) {
            const float inv = 1.0f / l;
#pragma unroll
            for (int j = 0; j < 8; ++j) o[j] *= inv;
            cnb_stv(out + pix * C + t.head * HD + sub * 8, o);
            if (sub == 0) lse[pix * g.heads + t.head] = m + logf(l);
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// backward, query side: dq_i (into the q third of dqkv) and pds[(pix*heads + head)*K2 + n] = (p_in, scale p_in (dp_in - D_i))
// ---------------------------------------------------------------------------------------------------------------------
template <int KS, int DIL, int LPH>
__global__ void __launch_bounds__(NA_TILE_THREADS) na2d_bwd_dq_fast_kernel(const bf16_t* __restrict__ qkv, const bf16_t* __restrict__ dout,
                                                                          const bf16_t* __restrict__ out, const float* __restrict__ lse,
                                                                          float2* __restrict__ pds, bf16_t* __restrict__ dqkv, NaTile g) {
    CNB_PDL_SYNC();
    constexpr int HD = LPH * 8, K2 = KS * KS;
    CNB_DYN_SMEM(sm_raw);
    bf16_t* sm = reinterpret_cast<bf16_t*>(sm_raw);
    const TilePos t = tile_pos(g);
    const int C = g.heads * HD;
    na_stage_region(sm, qkv + C + t.head * HD, qkv + 2 * C + t.head * HD, 3L * C, g, t.img_pix0, t.ry0, t.rx0);
    __syncthreads();
    const int dil = DIL > 0 ? DIL : g.dil;
    const int col_step = dil * 2 * HD, row_step = col_step * g.RW;

    for (int it = threadIdx.x; it < NA_TH * NA_TW * LPH; it += NA_TILE_THREADS) {
        const int sub = it % LPH, pl = it / LPH;
        const int lx = pl % NA_TW, ly = pl / NA_TW;
        const bool valid = (t.y0 + ly) < g.H && (t.x0 + lx) < g.W;
        const int y = (t.y0 + ly) < g.H ? t.y0 + ly : g.H - 1, x = (t.x0 + lx) < g.W ? t.x0 + lx : g.W - 1;
        const long pix = t.img_pix0 + (long)y * g.W + x;
        float q[8], go[8];
        unpack8(*reinterpret_cast<const uint4*>(qkv + pix * 3 * C + t.head * HD + sub * 8), q);
        const uint4 graw = *reinterpret_cast<const uint4*>(dout + pix * C + t.head * HD + sub * 8);
        unpack8(graw, go);
        const float D = gsum<LPH>(dot8(go, *reinterpret_cast<const uint4*>(out + pix * C + t.head * HD + sub * 8)));
        const float L = lse[pix * g.heads + t.head];
#pragma unroll
        for (int j = 0; j < 8; ++j) q[j] *= g.scale;
        const int sy = wstart<DIL>(y, g.H, KS, g.dil), sx = wstart<DIL>(x, g.W, KS, g.dil);
        const bf16_t* kb = sm + ((sy - t.ry0) * g.RW + (sx - t.rx0)) * 2 * HD + sub * 8;
        float dq[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) dq[j] = 0.f;
        float2* rec = pds + (pix * g.heads + t.head) * K2;
#pragma unroll
        for (int a = 0; a < KS; ++a)
#pragma unroll
            for (int b = 0; b < KS; ++b) {
                const uint4 kraw = *reinterpret_cast<const uint4*>(kb + a * row_step + b * col_step);
                const uint4 vraw = *reinterpret_cast<const uint4*>(kb + a * row_step + b * col_step + HD);
                const float s = gsum<LPH>(dot8(q, kraw)), dp = gsum<LPH>(dot8(go, vraw));
                const float p = cnb_exp(s - L);
                const float ds = p * (dp - D);
                axpy8(ds, kraw, dq);
                // the lanes of the group share p and ds: lane (n mod LPH) writes record n (72 contiguous bytes per group for k = 3)
                if (valid && sub == (a * KS + b) % LPH) rec[a * KS + b] = make_float2(p, ds * g.scale);
            }
        if (valid) {
#pragma unroll
            for (int j = 0; j < 8; ++j) dq[j] *= g.scale;
            cnb_stv(dqkv + pix * 3 * C + t.head * HD + sub * 8, dq);
        }
    }
}

// which query indices i = j + m*d (m in [-(k-1), k-1]) have j inside their clamped window: bit (m + k - 1)
template <int KS, int DIL>
__device__ __forceinline__ uint32_t inverse_mask(int j, int len, int dil_rt) {
    const int d = DIL > 0 ? DIL : dil_rt;
    uint32_t mask = 0;
#pragma unroll
    for (int m = -(KS - 1); m <= KS - 1; ++m) {
        const int i = j + m * d;
        if (i < 0 || i >= len) continue;
        const int s = wstart<DIL>(i, len, KS, dil_rt);
        if (s <= j && j <= s + (KS - 1) * d) mask |= 1u << (m + KS - 1);
    }
    return mask;
}

// ---------------------------------------------------------------------------------------------------------------------
// backward, key side: dk_j = sum_i ds_ij q_i, dv_j = sum_i p_ij dout_i over the queries i whose window holds j (a gather: no
// atomics, deterministic).  Region rows: q_i | dout_i.  A query clamped at the image border can sit up to (k-1)*d from its key,
// i.e. outside the staged region of an interior-side tile: those few candidates are read from global memory.
// ---------------------------------------------------------------------------------------------------------------------
template <int KS, int DIL, int LPH>
__global__ void __launch_bounds__(NA_TILE_THREADS) na2d_bwd_dkv_fast_kernel(const bf16_t* __restrict__ qkv, const bf16_t* __restrict__ dout,
                                                                           const float2* __restrict__ pds, bf16_t* __restrict__ dqkv,
                                                                           NaTile g) {
    CNB_PDL_SYNC();
    constexpr int HD = LPH * 8, K2 = KS * KS;
    CNB_DYN_SMEM(sm_raw);
    bf16_t* sm = reinterpret_cast<bf16_t*>(sm_raw);
    const TilePos t = tile_pos(g);
    const int C = g.heads * HD;
    {
        constexpr int parts = HD / 8;
        const int total = g.RH * g.RW * 2 * parts;
        for (int i = threadIdx.x; i < total; i += blockDim.x) {
            const int part = i % (2 * parts);
            const int r = i / (2 * parts);
            const int rx = r % g.RW, ry = r / g.RW;
            const long pix = t.img_pix0 + (long)(t.ry0 + ry) * g.W + (t.rx0 + rx);
            const bf16_t* src = part < parts ? qkv + pix * 3 * C + t.head * HD + part * 8 : dout + pix * C + t.head * HD + (part - parts) * 8;
            cnb_cp_async16(sm + (long)r * 2 * HD + part * 8, src);
        }
        cnb_cp_async_wait_all();
    }
    __syncthreads();
    const int dil = DIL > 0 ? DIL : g.dil;

    for (int it = threadIdx.x; it < NA_TH * NA_TW * LPH; it += NA_TILE_THREADS) {
        const int sub = it % LPH, pl = it / LPH;
        const int lx = pl % NA_TW, ly = pl / NA_TW;
        const int y = t.y0 + ly, x = t.x0 + lx;
        if (y >= g.H || x >= g.W) continue;  // no shuffles below: lanes may drop out
        const long pix = t.img_pix0 + (long)y * g.W + x;
        float dk[8], dv[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) dk[j] = 0.f, dv[j] = 0.f;
        const uint32_t ymask = inverse_mask<KS, DIL>(y, g.H, g.dil), xmask = inverse_mask<KS, DIL>(x, g.W, g.dil);
        // column index of x inside the window of each candidate query column (hoisted out of the row loop)
        int bcol[2 * KS - 1];
#pragma unroll
        for (int mx = 0; mx < 2 * KS - 1; ++mx) {
            const int ix = x + (mx - (KS - 1)) * dil;
            bcol[mx] = ((xmask >> mx) & 1u) ? (x - wstart<DIL>(ix, g.W, KS, g.dil)) / dil : 0;
        }
#pragma unroll
        for (int my = 0; my < 2 * KS - 1; ++my) {
            if (!((ymask >> my) & 1u)) continue;
            const int iy = y + (my - (KS - 1)) * dil;
            const int arow = (y - wstart<DIL>(iy, g.H, KS, g.dil)) / dil;
            const int ry = iy - t.ry0;
#pragma unroll
            for (int mx = 0; mx < 2 * KS - 1; ++mx) {
                if (!((xmask >> mx) & 1u)) continue;
                const int ix = x + (mx - (KS - 1)) * dil;
                const int rx = ix - t.rx0;
                const long ipix = t.img_pix0 + (long)iy * g.W + ix;
                const float2 w = pds[(ipix * g.heads + t.head) * K2 + arow * KS + bcol[mx]];
                uint4 qraw, graw;
                if (ry >= 0 && ry < g.RH && rx >= 0 && rx < g.RW) {
                    const bf16_t* qp = sm + (ry * g.RW + rx) * 2 * HD + sub * 8;
                    qraw = *reinterpret_cast<const uint4*>(qp);
                    graw = *reinterpret_cast<const uint4*>(qp + HD);
                } else {
                    qraw = *reinterpret_cast<const uint4*>(qkv + ipix * 3 * C + t.head * HD + sub * 8);
                    graw = *reinterpret_cast<const uint4*>(dout + ipix * C + t.head * HD + sub * 8);
                }
                axpy8(w.y, qraw, dk);
                axpy8(w.x, graw, dv);
            }
        }
        cnb_stv(dqkv + pix * 3 * C + C + t.head * HD + sub * 8, dk);
        cnb_stv(dqkv + pix * 3 * C + 2 * C + t.head * HD + sub * 8, dv);
    }
}

// shapes with a specialised instantiation: bf16, k in {3, 7}, dilation in {1, 2}, head_dim in {32, 64}
static inline bool eligible(int hd, int ksize, int dil, int dtype) {
    return dtype == CNB_BF16 && (ksize == 3 || ksize == 7) && (dil == 1 || dil == 2) && (hd == 32 || hd == 64);
}

}  // namespace naf
}  // namespace cnb
