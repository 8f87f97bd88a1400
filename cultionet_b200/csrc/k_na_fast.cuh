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
//   * every multiply-add takes its bf16 operands as the halves of packed words (fma.rn.f32.bf16 = FHFMA.BF16 on sm_100, fp32
//     accumulate): no bf16 -> fp32 unpacking, whose shift/mask pairs were half of the issued instructions; softmax probabilities
//     and logit gradients enter those products rounded to bf16 (as the P operand of a tensor-core attention kernel does);
//   * the backward no longer recomputes q.k and dout.v in the key-side pass: the query-side pass stores p_in and
//     scale * p_in (dp_in - D_i) as one bf16 pair (36 bytes per (pixel, head) for k = 3), and the key-side pass is a pure gather
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

// the same on packed operands (FHFMA.BF16, cnb_common.cuh): no unpacking, one instruction per element
__device__ __forceinline__ float dot8p(const uint4& a, const uint4& b) {
    float s = cnb_fma2_bf16(a.x, b.x, 0.f);
    s = cnb_fma2_bf16(a.y, b.y, s);
    s = cnb_fma2_bf16(a.z, b.z, s);
    return cnb_fma2_bf16(a.w, b.w, s);
}
// acc[0..8) += w * b, w = the low (HI = false) / high (HI = true) bf16 half of wp
template <bool HI>
__device__ __forceinline__ void axpy8p(uint32_t wp, const uint4& b, float* acc) {
    cnb_axpy2_bf16<HI>(wp, b.x, acc[0], acc[1]);
    cnb_axpy2_bf16<HI>(wp, b.y, acc[2], acc[3]);
    cnb_axpy2_bf16<HI>(wp, b.z, acc[4], acc[5]);
    cnb_axpy2_bf16<HI>(wp, b.w, acc[6], acc[7]);
}

// Dilation groups.  A pixel only ever attends to pixels of its own residue class (y mod d, x mod d), and natten clamps a window
// inside that class (oracle/natten_ref.py): neighbourhood attention with dilation d IS plain dilation-1 attention on each of the d*d
// sub-images x[gy::d, gx::d].  A CTA therefore works on a tile of ONE sub-image: the staged halo is k/2 sub-image pixels wide
// instead of (k/2)*d image pixels, i.e. the staged region of a k = 7, d = 2 tile shrinks from 20 x 28 (143 KB, one CTA per SM) to
// 14 x 22 pixels (79 KB, two CTAs per SM), and every kernel below is written for dilation 1 in sub-image coordinates
// (ys, xs) <-> image pixel (gy + d ys, gx + d xs) = img_pix0 + ys * rs + xs * cs.
struct TilePos {
    int head, y0, x0, ry0, rx0;
    int H, W, RH, RW;  // this sub-image, and its staged region
    long img_pix0, rs, cs;
    bool empty;  // the tile lies outside a (smaller) sub-image
};
__device__ __forceinline__ TilePos tile_pos(const NaTile& g) {
    TilePos t;
    t.head = blockIdx.y;
    const int d = g.groups > 1 ? g.groups : 1;
    int i = blockIdx.x;
    const int tx = i % g.tiles_x;
    i /= g.tiles_x;
    const int ty = i % g.tiles_y;
    i /= g.tiles_y;
    const int par = i % (d * d), b = i / (d * d);
    const int gy = par / d, gx = par % d;
    t.H = (g.H - gy + d - 1) / d, t.W = (g.W - gx + d - 1) / d;
    t.RH = g.RH < t.H ? g.RH : t.H, t.RW = g.RW < t.W ? g.RW : t.W;
    const int halo = g.ksize / 2;
    t.y0 = ty * NA_TH, t.x0 = tx * NA_TW;
    t.ry0 = na_region_origin(t.y0, halo, t.H, t.RH), t.rx0 = na_region_origin(t.x0, halo, t.W, t.RW);
    t.img_pix0 = (long)b * g.H * g.W + (long)gy * g.W + gx;
    t.rs = (long)d * g.W, t.cs = d;
    t.empty = t.y0 >= t.H || t.x0 >= t.W;
    return t;
}
// stage two HD-wide row segments per region pixel: smem[(ry*RW + rx)][0..HD) = a, [HD..2HD) = b
template <int HD>
__device__ __forceinline__ void stage_region(bf16_t* sm, const bf16_t* a_base, long a_stride, const bf16_t* b_base, long b_stride,
                                             const TilePos& t) {
    constexpr int parts = HD / 8;
    const int total = t.RH * t.RW * 2 * parts;
    for (int i = threadIdx.x; i < total; i += blockDim.x) {
        const int part = i % (2 * parts);
        const int r = i / (2 * parts);
        const int rx = r % t.RW, ry = r / t.RW;
        const long pix = t.img_pix0 + (t.ry0 + ry) * t.rs + (t.rx0 + rx) * t.cs;
        const bf16_t* src = part < parts ? a_base + pix * a_stride + part * 8 : b_base + pix * b_stride + (part - parts) * 8;
        cnb_cp_async16(sm + (long)r * 2 * HD + part * 8, src);
    }
    cnb_cp_async_wait_all();
}

// ---------------------------------------------------------------------------------------------------------------------
// forward: out_i = softmax_n(scale q_i . k_n) v_n, lse_i
// ---------------------------------------------------------------------------------------------------------------------
// VPL = 16-byte vectors per lane: a (pixel, head) is shared by LPH = HD / (8 VPL) lanes.  VPL = 2 for head_dim 64 halves the lanes per
// group: one shuffle step and half of the redundant exponentials / logit arithmetic per neighbour disappear (the kernels are
// issue-bound: ~27 instructions per lane and neighbour at VPL = 1, of which 3 shuffles + 3 adds + the exponential are per-lane
// overhead).  Lane `sub` of an even pixel holds channel chunks {sub, LPH + sub}, of an odd pixel {LPH + sub, sub}: the two groups of a
// quarter-warp then read different 64-byte halves of their 128-byte rows -- conflict-free 16-byte shared-memory loads.
template <int LPH, int VPL>
__device__ __forceinline__ void lane_chunks(int sub, int pl, int (&off)[VPL]) {
#pragma unroll
    for (int v = 0; v < VPL; ++v) off[v] = (((VPL > 1 ? (v ^ (pl & 1)) : v) * LPH) + sub) * 8;
}

template <int KS, int DIL, int LPH, int VPL>
__global__ void __launch_bounds__(NA_TILE_THREADS) na2d_fwd_fast_kernel(const bf16_t* __restrict__ qkv, bf16_t* __restrict__ out,
                                                                       float* __restrict__ lse, NaTile g) {
    CNB_PDL_SYNC();
    constexpr int HD = LPH * 8 * VPL, K2 = KS * KS;
    CNB_DYN_SMEM(sm_raw);
    bf16_t* sm = reinterpret_cast<bf16_t*>(sm_raw);
    static_assert(DIL == 1, "dilation is handled by the sub-image decomposition (tile_pos)");
    const TilePos t = tile_pos(g);
    if (t.empty) return;
    const int C = g.heads * HD;
    stage_region<HD>(sm, qkv + C + t.head * HD, 3L * C, qkv + 2 * C + t.head * HD, 3L * C, t);
    __syncthreads();
    const int col_step = 2 * HD, row_step = col_step * t.RW;  // elements between window columns / rows in the staged region

    static_assert((NA_TH * NA_TW * LPH) % NA_TILE_THREADS == 0, "whole rounds: the warps stay converged for the shuffles");
#pragma unroll 1
    for (int round = 0; round < (NA_TH * NA_TW * LPH) / NA_TILE_THREADS; ++round) {  // compile-time trip count: provably uniform
        const int it = round * NA_TILE_THREADS + threadIdx.x;
        const int sub = it % LPH, pl = it / LPH;
        const int lx = pl % NA_TW, ly = pl / NA_TW;
        const bool valid = (t.y0 + ly) < t.H && (t.x0 + lx) < t.W;
        // out-of-image lanes shadow the last pixel of the sub-image (it lies in this tile): control flow and shuffles stay uniform
        const int y = (t.y0 + ly) < t.H ? t.y0 + ly : t.H - 1, x = (t.x0 + lx) < t.W ? t.x0 + lx : t.W - 1;
        const long pix = t.img_pix0 + y * t.rs + x * t.cs;
        int off[VPL];
        lane_chunks<LPH, VPL>(sub, pl, off);
        uint4 qraw[VPL];
#pragma unroll
        for (int v = 0; v < VPL; ++v) qraw[v] = *reinterpret_cast<const uint4*>(qkv + pix * 3 * C + t.head * HD + off[v]);
        const int sy = wstart<1>(y, t.H, KS, 1), sx = wstart<1>(x, t.W, KS, 1);
        const bf16_t* kb = sm + ((sy - t.ry0) * t.RW + (sx - t.rx0)) * 2 * HD;
        float lg[K2];
        float m = -INFINITY;
#pragma unroll
        for (int a = 0; a < KS; ++a)
#pragma unroll
            for (int b = 0; b < KS; ++b) {
                float part = 0.f;
#pragma unroll
                for (int v = 0; v < VPL; ++v)
                    part += dot8p(qraw[v], *reinterpret_cast<const uint4*>(kb + a * row_step + b * col_step + off[v]));
                const float s = g.scale * gsum<LPH>(part);
                lg[a * KS + b] = s;
                m = fmaxf(m, s);
            }
        float l = 0.f;
#pragma unroll
        for (int n = 0; n < K2; ++n) {
            lg[n] = cnb_exp(lg[n] - m);
            l += lg[n];
        }
        float o[VPL][8];
#pragma unroll
        for (int v = 0; v < VPL; ++v)
#pragma unroll
            for (int j = 0; j < 8; ++j) o[v][j] = 0.f;
        // the un-normalised probabilities (in (0, 1]) weight v as bf16, like the P operand of a tensor-core attention kernel
#pragma unroll
        for (int a = 0; a < KS; ++a)
#pragma unroll
            for (int b = 0; b < KS; ++b) {
                const uint32_t w = cnb_pack_bf16x2(lg[a * KS + b], 0.f);
#pragma unroll
                for (int v = 0; v < VPL; ++v)
                    axpy8p<false>(w, *reinterpret_cast<const uint4*>(kb + a * row_step + b * col_step + HD + off[v]), o[v]);
            }
        if (valid) {
            const float inv = 1.0f / l;
#pragma unroll
            for (int v = 0; v < VPL; ++v) {
#pragma unroll
                for (int j = 0; j < 8; ++j) o[v][j] *= inv;
                cnb_stv(out + pix * C + t.head * HD + off[v], o[v]);
            }
            if (sub == 0) lse[pix * g.heads + t.head] = m + logf(l);
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// backward, query side: dq_i (into the q third of dqkv) and pds[(pix*heads + head)*K2 + n] = (p_in, scale p_in (dp_in - D_i))
// ---------------------------------------------------------------------------------------------------------------------
template <int KS, int DIL, int LPH, int VPL>
__global__ void __launch_bounds__(NA_TILE_THREADS) na2d_bwd_dq_fast_kernel(const bf16_t* __restrict__ qkv, const bf16_t* __restrict__ dout,
                                                                          const bf16_t* __restrict__ out, const float* __restrict__ lse,
                                                                          uint32_t* __restrict__ pds, bf16_t* __restrict__ dqkv, NaTile g) {
    CNB_PDL_SYNC();
    constexpr int HD = LPH * 8 * VPL, K2 = KS * KS;
    CNB_DYN_SMEM(sm_raw);
    bf16_t* sm = reinterpret_cast<bf16_t*>(sm_raw);
    static_assert(DIL == 1, "dilation is handled by the sub-image decomposition (tile_pos)");
    const TilePos t = tile_pos(g);
    if (t.empty) return;
    const int C = g.heads * HD;
    stage_region<HD>(sm, qkv + C + t.head * HD, 3L * C, qkv + 2 * C + t.head * HD, 3L * C, t);
    __syncthreads();
    const int col_step = 2 * HD, row_step = col_step * t.RW;

    static_assert((NA_TH * NA_TW * LPH) % NA_TILE_THREADS == 0, "whole rounds: the warps stay converged for the shuffles");
#pragma unroll 1
    for (int round = 0; round < (NA_TH * NA_TW * LPH) / NA_TILE_THREADS; ++round) {  // compile-time trip count: provably uniform
        const int it = round * NA_TILE_THREADS + threadIdx.x;
        const int sub = it % LPH, pl = it / LPH;
        const int lx = pl % NA_TW, ly = pl / NA_TW;
        const bool valid = (t.y0 + ly) < t.H && (t.x0 + lx) < t.W;
        const int y = (t.y0 + ly) < t.H ? t.y0 + ly : t.H - 1, x = (t.x0 + lx) < t.W ? t.x0 + lx : t.W - 1;
        const long pix = t.img_pix0 + y * t.rs + x * t.cs;
        int off[VPL];
        lane_chunks<LPH, VPL>(sub, pl, off);
        uint4 qraw[VPL], graw[VPL];
        float dpart = 0.f;
#pragma unroll
        for (int v = 0; v < VPL; ++v) {
            qraw[v] = *reinterpret_cast<const uint4*>(qkv + pix * 3 * C + t.head * HD + off[v]);
            graw[v] = *reinterpret_cast<const uint4*>(dout + pix * C + t.head * HD + off[v]);
            dpart += dot8p(graw[v], *reinterpret_cast<const uint4*>(out + pix * C + t.head * HD + off[v]));
        }
        const float D = gsum<LPH>(dpart);
        const float L = lse[pix * g.heads + t.head];
        const int sy = wstart<1>(y, t.H, KS, 1), sx = wstart<1>(x, t.W, KS, 1);
        const bf16_t* kb = sm + ((sy - t.ry0) * t.RW + (sx - t.rx0)) * 2 * HD;
        float dq[VPL][8];
#pragma unroll
        for (int v = 0; v < VPL; ++v)
#pragma unroll
            for (int j = 0; j < 8; ++j) dq[v][j] = 0.f;
        uint32_t* rec = pds + (pix * g.heads + t.head) * K2;
#pragma unroll
        for (int a = 0; a < KS; ++a)
#pragma unroll
            for (int b = 0; b < KS; ++b) {
                uint4 kraw[VPL];
                float sp = 0.f, dpp = 0.f;
#pragma unroll
                for (int v = 0; v < VPL; ++v) {
                    kraw[v] = *reinterpret_cast<const uint4*>(kb + a * row_step + b * col_step + off[v]);
                    const uint4 vraw = *reinterpret_cast<const uint4*>(kb + a * row_step + b * col_step + HD + off[v]);
                    sp += dot8p(qraw[v], kraw[v]);
                    dpp += dot8p(graw[v], vraw);
                }
                const float s = g.scale * gsum<LPH>(sp), dp = gsum<LPH>(dpp);
                const float p = cnb_exp(s - L);
                const float ds = p * (dp - D) * g.scale;
                // record = (p, scale * ds) as one bf16 pair: the weights of this pass (ds, high half) and of the key-side pass
                const uint32_t w = cnb_pack_bf16x2(p, ds);
#pragma unroll
                for (int v = 0; v < VPL; ++v) axpy8p<true>(w, kraw[v], dq[v]);
                // the lanes of the group share the record: lane (n mod LPH) writes record n (contiguous bytes per group)
                if (valid && sub == (a * KS + b) % LPH) rec[a * KS + b] = w;
            }
        if (valid) {
#pragma unroll
            for (int v = 0; v < VPL; ++v) cnb_stv(dqkv + pix * 3 * C + t.head * HD + off[v], dq[v]);  // ds already carries the scale
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
                                                                           const uint32_t* __restrict__ pds, bf16_t* __restrict__ dqkv,
                                                                           NaTile g) {
    CNB_PDL_SYNC();
    constexpr int HD = LPH * 8, K2 = KS * KS;
    CNB_DYN_SMEM(sm_raw);
    bf16_t* sm = reinterpret_cast<bf16_t*>(sm_raw);
    static_assert(DIL == 1, "dilation is handled by the sub-image decomposition (tile_pos)");
    const TilePos t = tile_pos(g);
    if (t.empty) return;
    const int C = g.heads * HD;
    stage_region<HD>(sm, qkv + t.head * HD, 3L * C, dout + t.head * HD, (long)C, t);
    __syncthreads();
    constexpr int dil = 1;

    for (int it = threadIdx.x; it < NA_TH * NA_TW * LPH; it += NA_TILE_THREADS) {
        const int sub = it % LPH, pl = it / LPH;
        const int lx = pl % NA_TW, ly = pl / NA_TW;
        const int y = t.y0 + ly, x = t.x0 + lx;
        if (y >= t.H || x >= t.W) continue;  // no shuffles below: lanes may drop out
        const long pix = t.img_pix0 + y * t.rs + x * t.cs;
        float dk[8], dv[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) dk[j] = 0.f, dv[j] = 0.f;
        // Interior keys (most of the image): no window that holds the key is clamped, so the queries are exactly the k x k pixels
        // (y + my*d, x + mx*d), |my|, |mx| <= k/2, the key sits at window position (k/2 - my, k/2 - mx) of each, and all of them lie
        // in the staged region: compile-time offsets, no window arithmetic (the generic path below spends ~60 integer instructions
        // per candidate on it, 5 x 5 candidates for k = 3).  A clamped window reaches at most (k - 1)*d + d - 1 from its border.
        constexpr int HK = KS / 2;
        const int lim = (2 * HK + 1) * dil;
        if (y >= lim && y + lim < t.H && x >= lim && x + lim < t.W) {
            const bf16_t* qb = sm + ((y - t.ry0) * t.RW + (x - t.rx0)) * 2 * HD + sub * 8;
            const uint32_t* rb = pds + (pix * g.heads + t.head) * K2;
            const long rec_row = t.rs * g.heads * K2, rec_col = t.cs * g.heads * K2;
            const int sm_row = t.RW * 2 * HD, sm_col = 2 * HD;
#pragma unroll
            for (int my = -HK; my <= HK; ++my)
#pragma unroll
                for (int mx = -HK; mx <= HK; ++mx) {
                    const uint32_t w = rb[my * rec_row + mx * rec_col + (HK - my) * KS + (HK - mx)];
                    const bf16_t* qp = qb + my * sm_row + mx * sm_col;
                    axpy8p<true>(w, *reinterpret_cast<const uint4*>(qp), dk);
                    axpy8p<false>(w, *reinterpret_cast<const uint4*>(qp + HD), dv);
                }
            cnb_stv(dqkv + pix * 3 * C + C + t.head * HD + sub * 8, dk);
            cnb_stv(dqkv + pix * 3 * C + 2 * C + t.head * HD + sub * 8, dv);
            continue;
        }
        const uint32_t ymask = inverse_mask<KS, 1>(y, t.H, 1), xmask = inverse_mask<KS, 1>(x, t.W, 1);
        // column index of x inside the window of each candidate query column (hoisted out of the row loop)
        int bcol[2 * KS - 1];
#pragma unroll
        for (int mx = 0; mx < 2 * KS - 1; ++mx) {
            const int ix = x + (mx - (KS - 1)) * dil;
            bcol[mx] = ((xmask >> mx) & 1u) ? (x - wstart<1>(ix, t.W, KS, 1)) : 0;
        }
#pragma unroll
        for (int my = 0; my < 2 * KS - 1; ++my) {
            if (!((ymask >> my) & 1u)) continue;
            const int iy = y + (my - (KS - 1)) * dil;
            const int arow = y - wstart<1>(iy, t.H, KS, 1);
            const int ry = iy - t.ry0;
#pragma unroll
            for (int mx = 0; mx < 2 * KS - 1; ++mx) {
                if (!((xmask >> mx) & 1u)) continue;
                const int ix = x + (mx - (KS - 1)) * dil;
                const int rx = ix - t.rx0;
                const long ipix = t.img_pix0 + iy * t.rs + ix * t.cs;
                const uint32_t w = pds[(ipix * g.heads + t.head) * K2 + arow * KS + bcol[mx]];  // bf16 pair (p, scale * ds)
                uint4 qraw, graw;
                if (ry >= 0 && ry < t.RH && rx >= 0 && rx < t.RW) {
                    const bf16_t* qp = sm + (ry * t.RW + rx) * 2 * HD + sub * 8;
                    qraw = *reinterpret_cast<const uint4*>(qp);
                    graw = *reinterpret_cast<const uint4*>(qp + HD);
                } else {
                    qraw = *reinterpret_cast<const uint4*>(qkv + ipix * 3 * C + t.head * HD + sub * 8);
                    graw = *reinterpret_cast<const uint4*>(dout + ipix * C + t.head * HD + sub * 8);
                }
                axpy8p<true>(w, qraw, dk);
                axpy8p<false>(w, graw, dv);
            }
        }
        cnb_stv(dqkv + pix * 3 * C + C + t.head * HD + sub * 8, dk);
        cnb_stv(dqkv + pix * 3 * C + 2 * C + t.head * HD + sub * 8, dv);
    }
}

struct TilePosImg {
    int head, b, y0, x0, ry0, rx0;
    long img_pix0;
};
__device__ __forceinline__ TilePosImg tile_pos_img(const NaTile& g) {
    TilePosImg t;
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
// backward, key side, IMAGE-SPACE tiles (the version before the sub-image decomposition; NaTile from na_tile_setup, dilation as a
// template parameter).  Kept selectable (CNB_NA_DKV=img) because the key-side gather of the query records behaves differently in L1
// under the two tilings.
// backward, key side: dk_j = sum_i ds_ij q_i, dv_j = sum_i p_ij dout_i over the queries i whose window holds j (a gather: no
// atomics, deterministic).  Region rows: q_i | dout_i.  A query clamped at the image border can sit up to (k-1)*d from its key,
// i.e. outside the staged region of an interior-side tile: those few candidates are read from global memory.
// ---------------------------------------------------------------------------------------------------------------------
template <int KS, int DIL, int LPH, int VPL>
__global__ void __launch_bounds__(NA_TILE_THREADS) na2d_bwd_dkv_img_kernel(const bf16_t* __restrict__ qkv, const bf16_t* __restrict__ dout,
                                                                           const uint32_t* __restrict__ pds, bf16_t* __restrict__ dqkv,
                                                                           NaTile g) {
    CNB_PDL_SYNC();
    constexpr int HD = LPH * 8 * VPL, K2 = KS * KS;
    CNB_DYN_SMEM(sm_raw);
    bf16_t* sm = reinterpret_cast<bf16_t*>(sm_raw);
    const TilePosImg t = tile_pos_img(g);
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
        int off[VPL];
        lane_chunks<LPH, VPL>(sub, pl, off);
        float dk[VPL][8], dv[VPL][8];
#pragma unroll
        for (int v = 0; v < VPL; ++v)
#pragma unroll
            for (int j = 0; j < 8; ++j) dk[v][j] = 0.f, dv[v][j] = 0.f;
        // Interior keys (most of the image): no window that holds the key is clamped, so the queries are exactly the k x k pixels
        // (y + my*d, x + mx*d), |my|, |mx| <= k/2, the key sits at window position (k/2 - my, k/2 - mx) of each, and all of them lie
        // in the staged region: compile-time offsets, no window arithmetic (the generic path below spends ~60 integer instructions
        // per candidate on it, 5 x 5 candidates for k = 3).  A clamped window reaches at most (k - 1)*d + d - 1 from its border.
        constexpr int HK = KS / 2;
        const int lim = (2 * HK + 1) * dil;
        if (y >= lim && y + lim < g.H && x >= lim && x + lim < g.W) {
            const bf16_t* qb = sm + ((y - t.ry0) * g.RW + (x - t.rx0)) * 2 * HD;
            const uint32_t* rb = pds + (pix * g.heads + t.head) * K2;
            const long rec_row = (long)dil * g.W * g.heads * K2, rec_col = (long)dil * g.heads * K2;
            const int sm_row = dil * g.RW * 2 * HD, sm_col = dil * 2 * HD;
#pragma unroll
            for (int my = -HK; my <= HK; ++my)
#pragma unroll
                for (int mx = -HK; mx <= HK; ++mx) {
                    const uint32_t w = rb[my * rec_row + mx * rec_col + (HK - my) * KS + (HK - mx)];
                    const bf16_t* qp = qb + my * sm_row + mx * sm_col;
#pragma unroll
                    for (int v = 0; v < VPL; ++v) {
                        axpy8p<true>(w, *reinterpret_cast<const uint4*>(qp + off[v]), dk[v]);
                        axpy8p<false>(w, *reinterpret_cast<const uint4*>(qp + HD + off[v]), dv[v]);
                    }
                }
#pragma unroll
            for (int v = 0; v < VPL; ++v) {
                cnb_stv(dqkv + pix * 3 * C + C + t.head * HD + off[v], dk[v]);
                cnb_stv(dqkv + pix * 3 * C + 2 * C + t.head * HD + off[v], dv[v]);
            }
            continue;
        }
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
                const uint32_t w = pds[(ipix * g.heads + t.head) * K2 + arow * KS + bcol[mx]];  // bf16 pair (p, scale * ds)
                const bool staged = ry >= 0 && ry < g.RH && rx >= 0 && rx < g.RW;
#pragma unroll
                for (int v = 0; v < VPL; ++v) {
                    uint4 qraw, graw;
                    if (staged) {
                        const bf16_t* qp = sm + (ry * g.RW + rx) * 2 * HD + off[v];
                        qraw = *reinterpret_cast<const uint4*>(qp);
                        graw = *reinterpret_cast<const uint4*>(qp + HD);
                    } else {
                        qraw = *reinterpret_cast<const uint4*>(qkv + ipix * 3 * C + t.head * HD + off[v]);
                        graw = *reinterpret_cast<const uint4*>(dout + ipix * C + t.head * HD + off[v]);
                    }
                    axpy8p<true>(w, qraw, dk[v]);
                    axpy8p<false>(w, graw, dv[v]);
                }
            }
        }
#pragma unroll
        for (int v = 0; v < VPL; ++v) {
            cnb_stv(dqkv + pix * 3 * C + C + t.head * HD + off[v], dk[v]);
            cnb_stv(dqkv + pix * 3 * C + 2 * C + t.head * HD + off[v], dv[v]);
        }
    }
}

// shapes with a specialised instantiation: bf16, k in {3, 7}, dilation in {1, 2}, head_dim in {32, 64}
static inline bool eligible(int hd, int ksize, int dil, int dtype) {
    return dtype == CNB_BF16 && (ksize == 3 || ksize == 7) && (dil == 1 || dil == 2) && (hd == 32 || hd == 64);
}

}  // namespace naf
}  // namespace cnb
