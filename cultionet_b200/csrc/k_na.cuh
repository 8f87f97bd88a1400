// 2-D neighbourhood attention core (QK^T over a clamped dilated k x k window, softmax, PV).
// One warp per (pixel, head); lanes stride the head dimension, logits live in registers and the
// softmax is a warp-shuffle reduction.  Window rule: oracle/natten_ref.py::window_start.
#pragma once
#include "cnb_common.cuh"

namespace cnb {

constexpr int NA_MAX_K = 9;
constexpr int NA_MAX_K2 = NA_MAX_K * NA_MAX_K;
constexpr int NA_MAX_DPL = 4;  // head_dim <= 128

__device__ __forceinline__ int na_window_start(int index, int length, int ksize, int dilation) {
    const int g = index % dilation;
    const int p = index / dilation;
    const int group_len = (length - g + dilation - 1) / dilation;
    int s = p - ksize / 2;
    if (s < 0) s = 0;
    if (s > group_len - ksize) s = group_len - ksize;
    return g + dilation * s;
}

// Attention dropout (natten attn_drop, reference convolution.py:341-350): `rng` = device {seed, step counter} or null; probability n of
// item (pixel, head) is kept iff its 16 random bits >= drop_thr and then scaled by drop_scale = 1 / (1 - p).
struct NaDrop {
    const int64_t* rng;
    int site;
    uint32_t thr;
    float scale;
};

template <typename T>
__global__ void __launch_bounds__(256) na2d_fwd_kernel(const T* __restrict__ qkv, T* __restrict__ out, int B, int H, int W, int heads,
                                                      int hd, int ksize, int dil, float scale, NaDrop drop) {
    CNB_PDL_SYNC();
    const uint64_t drop_key = drop.rng ? cnb_rng_key(drop.rng, drop.site) : 0;
    const int lane = threadIdx.x & 31;
    const long warp = ((long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long nwarps = ((long)gridDim.x * blockDim.x) >> 5;
    const int C = heads * hd;
    const long total = (long)B * H * W * heads;
    const int k2 = ksize * ksize;
    for (long item = warp; item < total; item += nwarps) {
        const int head = (int)(item % heads);
        const long pix = item / heads;
        const int x = (int)(pix % W);
        const int y = (int)((pix / W) % H);
        const long img = pix / ((long)W * H);
        const int sy = na_window_start(y, H, ksize, dil);
        const int sx = na_window_start(x, W, ksize, dil);
        const T* qp = qkv + pix * 3 * C + head * hd;
        float q[NA_MAX_DPL];
#pragma unroll
        for (int j = 0; j < NA_MAX_DPL; ++j) {
            const int dd = lane + 32 * j;
            q[j] = dd < hd ? cnb_ld(qp + dd) * scale : 0.f;
        }
        float logit[NA_MAX_K2];
        float mx = -INFINITY;
        for (int n = 0; n < k2; ++n) {
            const int a = n / ksize, b = n - a * ksize;
            const long np = (img * H + (sy + a * dil)) * W + (sx + b * dil);
            const T* kp = qkv + np * 3 * C + C + head * hd;
            float part = 0.f;
#pragma unroll
            for (int j = 0; j < NA_MAX_DPL; ++j) {
                const int dd = lane + 32 * j;
                if (dd < hd) part = fmaf(q[j], cnb_ld(kp + dd), part);
            }
            const float l = cnb_warp_sum(part);
            logit[n] = l;
            mx = fmaxf(mx, l);
        }
        float den = 0.f;
        for (int n = 0; n < k2; ++n) {
            const float e = cnb_exp(logit[n] - mx);
            logit[n] = e;
            den += e;
        }
        const float inv = 1.0f / den;
        float o[NA_MAX_DPL];
#pragma unroll
        for (int j = 0; j < NA_MAX_DPL; ++j) o[j] = 0.f;
        for (int n = 0; n < k2; ++n) {
            const int a = n / ksize, b = n - a * ksize;
            const long np = (img * H + (sy + a * dil)) * W + (sx + b * dil);
            const T* vp = qkv + np * 3 * C + 2 * C + head * hd;
            float p = logit[n] * inv;
            if (drop.rng) p = cnb_rng_bits16(drop_key, (uint64_t)item * k2 + n) >= drop.thr ? p * drop.scale : 0.f;
#pragma unroll
            for (int j = 0; j < NA_MAX_DPL; ++j) {
                const int dd = lane + 32 * j;
                if (dd < hd) o[j] = fmaf(p, cnb_ld(vp + dd), o[j]);
            }
        }
        T* op = out + pix * C + head * hd;
#pragma unroll
        for (int j = 0; j < NA_MAX_DPL; ++j) {
            const int dd = lane + 32 * j;
            if (dd < hd) cnb_st(op + dd, o[j]);
        }
    }
}

// Recomputes the attention probabilities; dq is written, dk/dv are scattered with fp32 atomics into dacc.
template <typename T>
__global__ void __launch_bounds__(256) na2d_bwd_kernel(const T* __restrict__ qkv, const T* __restrict__ dout, float* __restrict__ dacc,
                                                      int B, int H, int W, int heads, int hd, int ksize, int dil, float scale, NaDrop drop) {
    CNB_PDL_SYNC();
    const uint64_t drop_key = drop.rng ? cnb_rng_key(drop.rng, drop.site) : 0;
    const int lane = threadIdx.x & 31;
    const long warp = ((long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long nwarps = ((long)gridDim.x * blockDim.x) >> 5;
    const int C = heads * hd;
    const long total = (long)B * H * W * heads;
    const int k2 = ksize * ksize;
    for (long item = warp; item < total; item += nwarps) {
        const int head = (int)(item % heads);
        const long pix = item / heads;
        const int x = (int)(pix % W);
        const int y = (int)((pix / W) % H);
        const long img = pix / ((long)W * H);
        const int sy = na_window_start(y, H, ksize, dil);
        const int sx = na_window_start(x, W, ksize, dil);
        const T* qp = qkv + pix * 3 * C + head * hd;
        const T* dop = dout + pix * C + head * hd;
        float q[NA_MAX_DPL], dq[NA_MAX_DPL], go[NA_MAX_DPL];
#pragma unroll
        for (int j = 0; j < NA_MAX_DPL; ++j) {
            const int dd = lane + 32 * j;
            q[j] = dd < hd ? cnb_ld(qp + dd) * scale : 0.f;
            go[j] = dd < hd ? cnb_ld(dop + dd) : 0.f;
            dq[j] = 0.f;
        }
        float prob[NA_MAX_K2], dprob[NA_MAX_K2];
        float mx = -INFINITY;
        for (int n = 0; n < k2; ++n) {
            const int a = n / ksize, b = n - a * ksize;
            const long np = (img * H + (sy + a * dil)) * W + (sx + b * dil);
            const T* kp = qkv + np * 3 * C + C + head * hd;
            const T* vp = kp + C;
            float part = 0.f, dpart = 0.f;
#pragma unroll
            for (int j = 0; j < NA_MAX_DPL; ++j) {
                const int dd = lane + 32 * j;
                if (dd < hd) {
                    part = fmaf(q[j], cnb_ld(kp + dd), part);
                    dpart = fmaf(go[j], cnb_ld(vp + dd), dpart);
                }
            }
            const float l = cnb_warp_sum(part);
            prob[n] = l;
            dprob[n] = cnb_warp_sum(dpart);
            mx = fmaxf(mx, l);
        }
        float den = 0.f;
        for (int n = 0; n < k2; ++n) {
            const float e = cnb_exp(prob[n] - mx);
            prob[n] = e;
            den += e;
        }
        const float inv = 1.0f / den;
        float dot = 0.f;
        for (int n = 0; n < k2; ++n) {
            prob[n] *= inv;
            // with dropout out = sum_n m_n p_n v_n (m_n = keep / (1 - p_drop)): d out / d p_n carries the same factor
            if (drop.rng) dprob[n] *= cnb_rng_bits16(drop_key, (uint64_t)item * k2 + n) >= drop.thr ? drop.scale : 0.f;
            dot = fmaf(prob[n], dprob[n], dot);
        }
        for (int n = 0; n < k2; ++n) {
            const int a = n / ksize, b = n - a * ksize;
            const long np = (img * H + (sy + a * dil)) * W + (sx + b * dil);
            const T* kp = qkv + np * 3 * C + C + head * hd;
            float* dkp = dacc + np * 3 * C + C + head * hd;
            float* dvp = dkp + C;
            const float p = prob[n];
            const float ds = p * (dprob[n] - dot);
            float pv = p;  // the (dropped, rescaled) probability that multiplied v_n in the forward
            if (drop.rng) pv = cnb_rng_bits16(drop_key, (uint64_t)item * k2 + n) >= drop.thr ? p * drop.scale : 0.f;
#pragma unroll
            for (int j = 0; j < NA_MAX_DPL; ++j) {
                const int dd = lane + 32 * j;
                if (dd < hd) {
                    dq[j] = fmaf(ds, cnb_ld(kp + dd), dq[j]);
                    atomicAdd(dkp + dd, ds * q[j]);
                    atomicAdd(dvp + dd, pv * go[j]);
                }
            }
        }
        float* dqp = dacc + pix * 3 * C + head * hd;
#pragma unroll
        for (int j = 0; j < NA_MAX_DPL; ++j) {
            const int dd = lane + 32 * j;
            if (dd < hd) dqp[dd] = dq[j] * scale;
        }
    }
}

template <typename T>
__global__ void __launch_bounds__(256) cast_from_f32_kernel(const float* __restrict__ src, T* __restrict__ dst, long n) {
    CNB_PDL_SYNC();
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) cnb_st(dst + i, src[i]);
}


// =====================================================================================================================
// Tiled kernels: one CTA = an 8 x 16 pixel tile of ONE head.  The k/v (forward, dq pass) or q/dout (dk/dv pass) rows of
// the tile plus its halo are staged once in shared memory with 128-bit loads, so HBM/L2 sees every operand ~once instead
// of k*k times.  hd/V lanes cooperate on one (pixel, head): each owns V channels (16 bytes) and the dot products are
// finished with xor shuffles inside that lane group.  The softmax is evaluated online (running max / sum), so no k*k
// register array and any odd kernel size works.  The backward is in gather form (no atomics, deterministic):
//   pass A (per query i):  dq_i = scale * sum_n p_in (dp_in - D_i) k_n,   D_i = dout_i . out_i,  p_in = exp(s_in - lse_i)
//   pass B (per key j):    dk_j = sum_{i: j in N(i)} p_ij (dp_ij - D_i) scale q_i,   dv_j = sum_i p_ij dout_i
// Halo: the clamped window of a pixel always contains the pixel and spans (k-1)*d, and a window clamped at a border lies
// within k*d of that border, so a region of TILE + (k-1)*d per axis, slid to stay inside the image, covers every window
// of the tile as long as TILE >= d (proof in DESIGN.md 4.3).
// =====================================================================================================================
constexpr int NA_TH = 8, NA_TW = 16;
constexpr int NA_TILE_THREADS = 256;
constexpr int NA_MAX_SMEM = 200 * 1024;

struct NaTile {
    int B, H, W, heads, hd, ksize, dil;
    int tiles_x, tiles_y, RH, RW;  // region = tile + halo, clipped to the image
    float scale;
    int groups;  // k_na_fast.cuh: dilation d handled as d*d independent dilation-1 sub-images (groups = d; H/W/tiles/RH/RW as set up
                 // by naf_tile_setup describe the LARGEST sub-image); 0 or 1 elsewhere
};

__device__ __forceinline__ int na_region_origin(int t0, int halo, int len, int rlen) {
    int o = t0 - halo;
    if (o > len - rlen) o = len - rlen;
    if (o < 0) o = 0;
    return o;
}

// sum over the LPH lanes that share one (pixel, head).  The shuffle names only that lane group: near the image border different
// pixels of a warp walk different neighbour lists, so the warp as a whole is NOT converged here.
template <int LPH>
__device__ __forceinline__ float na_group_sum(float v) {
    const unsigned gmask = LPH == 32 ? 0xffffffffu : (((1u << (LPH & 31)) - 1u) << ((threadIdx.x & 31u) & ~(unsigned)(LPH - 1)));
#pragma unroll
    for (int o = LPH / 2; o > 0; o >>= 1) v += __shfl_xor_sync(gmask, v, o);
    return v;
}

// stage two hd-wide row segments per region pixel: smem[(ry*RW + rx)][0..hd) = a, [hd..2hd) = b
template <typename T>
__device__ __forceinline__ void na_stage_region(T* sm, const T* a_base, const T* b_base, long pix_stride, const NaTile& g, long img_pix0,
                                                int ry0, int rx0) {
    constexpr int V = cnb_vec<T>::N;
    const int parts = g.hd / V;
    const int total = g.RH * g.RW * 2 * parts;
    for (int i = threadIdx.x; i < total; i += blockDim.x) {
        const int part = i % (2 * parts);
        const int r = i / (2 * parts);
        const int rx = r % g.RW, ry = r / g.RW;
        const long pix = img_pix0 + (long)(ry0 + ry) * g.W + (rx0 + rx);
        const T* src = (part < parts ? a_base + part * V : b_base + (part - parts) * V) + pix * pix_stride;
        cnb_cp_async16(sm + (long)r * 2 * g.hd + part * V, src);
    }
    cnb_cp_async_wait_all();
}

// k / v row segment of neighbour (ny, nx): from the staged region, or straight from global memory should a neighbour ever fall
// outside it (the region is sized so that it does not; this keeps correctness independent of that argument)
template <typename T>
__device__ __forceinline__ void na_kv_ptr(const T* sm, const T* qkv, const NaTile& g, int C, int head, int choff, long img_pix0, int ry0,
                                          int rx0, int ny, int nx, const T*& kp, const T*& vp) {
    const int ry = ny - ry0, rx = nx - rx0;
    if (ry >= 0 && ry < g.RH && rx >= 0 && rx < g.RW) {
        kp = sm + ((long)ry * g.RW + rx) * 2 * g.hd + choff;
        vp = kp + g.hd;
    } else {
        kp = qkv + (img_pix0 + (long)ny * g.W + nx) * 3 * C + C + head * g.hd + choff;
        vp = kp + C;
    }
}

template <typename T, int LPH>
__global__ void __launch_bounds__(NA_TILE_THREADS) na2d_fwd_tile_kernel(const T* __restrict__ qkv, T* __restrict__ out,
                                                                       float* __restrict__ lse, NaTile g) {
    CNB_PDL_SYNC();
    constexpr int V = cnb_vec<T>::N;
    CNB_DYN_SMEM(sm_raw);
    T* sm = reinterpret_cast<T*>(sm_raw);
    const int head = blockIdx.y;
    int t = blockIdx.x;
    const int tx = t % g.tiles_x;
    t /= g.tiles_x;
    const int ty = t % g.tiles_y;
    const int b = t / g.tiles_y;
    const int C = g.heads * g.hd;
    const int halo = (g.ksize / 2) * g.dil;
    const int y0 = ty * NA_TH, x0 = tx * NA_TW;
    const int ry0 = na_region_origin(y0, halo, g.H, g.RH), rx0 = na_region_origin(x0, halo, g.W, g.RW);
    const long img_pix0 = (long)b * g.H * g.W;
    na_stage_region(sm, qkv + C + head * g.hd, qkv + 2 * C + head * g.hd, 3L * C, g, img_pix0, ry0, rx0);
    __syncthreads();

    const int items = NA_TH * NA_TW * LPH;
    for (int it = threadIdx.x; it < items; it += blockDim.x) {
        const int sub = it % LPH;
        const int pl = it / LPH;
        const int lx = pl % NA_TW, ly = pl / NA_TW;
        const bool valid = (y0 + ly) < g.H && (x0 + lx) < g.W;
        // out-of-image lanes shadow a real pixel OF THIS TILE (shuffles stay uniform, the staged region covers its window)
        const int y = (y0 + ly) < g.H ? y0 + ly : g.H - 1, x = (x0 + lx) < g.W ? x0 + lx : g.W - 1;
        const long pix = img_pix0 + (long)y * g.W + x;
        float q[V];
        cnb_ldv(qkv + pix * 3 * C + head * g.hd + sub * V, q);
#pragma unroll
        for (int j = 0; j < V; ++j) q[j] *= g.scale;
        const int sy = na_window_start(y, g.H, g.ksize, g.dil), sx = na_window_start(x, g.W, g.ksize, g.dil);
        float m = -INFINITY, l = 0.f, o[V];
#pragma unroll
        for (int j = 0; j < V; ++j) o[j] = 0.f;
        for (int a = 0; a < g.ksize; ++a) {
            const int ny = sy + a * g.dil;
            for (int bb = 0; bb < g.ksize; ++bb) {
                const int nx = sx + bb * g.dil;
                const T *kp, *vp;
                na_kv_ptr(sm, qkv, g, C, head, sub * V, img_pix0, ry0, rx0, ny, nx, kp, vp);
                float kv[V];
                cnb_ldv(kp, kv);
                float part = 0.f;
#pragma unroll
                for (int j = 0; j < V; ++j) part = fmaf(q[j], kv[j], part);
                const float sc = na_group_sum<LPH>(part);
                const float mn = fmaxf(m, sc);
                const float corr = cnb_exp(m - mn), pe = cnb_exp(sc - mn);
                cnb_ldv(vp, kv);
                l = fmaf(l, corr, pe);
#pragma unroll
                for (int j = 0; j < V; ++j) o[j] = fmaf(o[j], corr, pe * kv[j]);
                m = mn;
            }
        }
        if (valid) {
            const float inv = 1.0f / l;
#pragma unroll
            for (int j = 0; j < V; ++j) o[j] *= inv;
            cnb_stv(out + pix * C + head * g.hd + sub * V, o);
            if (sub == 0) lse[pix * g.heads + head] = m + logf(l);
        }
    }
}

// pass A: dq (written into the q third of dqkv) and D_i
template <typename T, int LPH>
__global__ void __launch_bounds__(NA_TILE_THREADS) na2d_bwd_dq_tile_kernel(const T* __restrict__ qkv, const T* __restrict__ dout,
                                                                          const T* __restrict__ out, const float* __restrict__ lse,
                                                                          float* __restrict__ dvec, T* __restrict__ dqkv, NaTile g) {
    CNB_PDL_SYNC();
    constexpr int V = cnb_vec<T>::N;
    CNB_DYN_SMEM(sm_raw);
    T* sm = reinterpret_cast<T*>(sm_raw);
    const int head = blockIdx.y;
    int t = blockIdx.x;
    const int tx = t % g.tiles_x;
    t /= g.tiles_x;
    const int ty = t % g.tiles_y;
    const int b = t / g.tiles_y;
    const int C = g.heads * g.hd;
    const int halo = (g.ksize / 2) * g.dil;
    const int y0 = ty * NA_TH, x0 = tx * NA_TW;
    const int ry0 = na_region_origin(y0, halo, g.H, g.RH), rx0 = na_region_origin(x0, halo, g.W, g.RW);
    const long img_pix0 = (long)b * g.H * g.W;
    na_stage_region(sm, qkv + C + head * g.hd, qkv + 2 * C + head * g.hd, 3L * C, g, img_pix0, ry0, rx0);
    __syncthreads();

    const int items = NA_TH * NA_TW * LPH;
    for (int it = threadIdx.x; it < items; it += blockDim.x) {
        const int sub = it % LPH;
        const int pl = it / LPH;
        const int lx = pl % NA_TW, ly = pl / NA_TW;
        const bool valid = (y0 + ly) < g.H && (x0 + lx) < g.W;
        const int y = (y0 + ly) < g.H ? y0 + ly : g.H - 1, x = (x0 + lx) < g.W ? x0 + lx : g.W - 1;
        const long pix = img_pix0 + (long)y * g.W + x;
        float q[V], go[V], dq[V];
        cnb_ldv(qkv + pix * 3 * C + head * g.hd + sub * V, q);
        cnb_ldv(dout + pix * C + head * g.hd + sub * V, go);
        float part = 0.f;
        {
            float ov[V];
            cnb_ldv(out + pix * C + head * g.hd + sub * V, ov);
#pragma unroll
            for (int j = 0; j < V; ++j) part = fmaf(go[j], ov[j], part);
        }
        const float D = na_group_sum<LPH>(part);
        const float L = lse[pix * g.heads + head];
#pragma unroll
        for (int j = 0; j < V; ++j) {
            q[j] *= g.scale;
            dq[j] = 0.f;
        }
        const int sy = na_window_start(y, g.H, g.ksize, g.dil), sx = na_window_start(x, g.W, g.ksize, g.dil);
        for (int a = 0; a < g.ksize; ++a) {
            const int ny = sy + a * g.dil;
            for (int bb = 0; bb < g.ksize; ++bb) {
                const int nx = sx + bb * g.dil;
                const T *kp, *vp;
                na_kv_ptr(sm, qkv, g, C, head, sub * V, img_pix0, ry0, rx0, ny, nx, kp, vp);
                float kv[V], vv[V];
                cnb_ldv(kp, kv);
                cnb_ldv(vp, vv);
                float ps = 0.f, pd = 0.f;
#pragma unroll
                for (int j = 0; j < V; ++j) {
                    ps = fmaf(q[j], kv[j], ps);
                    pd = fmaf(go[j], vv[j], pd);
                }
                const float sc = na_group_sum<LPH>(ps), dp = na_group_sum<LPH>(pd);
                const float ds = cnb_exp(sc - L) * (dp - D);
#pragma unroll
                for (int j = 0; j < V; ++j) dq[j] = fmaf(ds, kv[j], dq[j]);
            }
        }
        if (valid) {
#pragma unroll
            for (int j = 0; j < V; ++j) dq[j] *= g.scale;
            cnb_stv(dqkv + pix * 3 * C + head * g.hd + sub * V, dq);
            if (sub == 0) dvec[pix * g.heads + head] = D;
        }
    }
}

// which query indices i (same dilation group, i = j + m*d, m in [-(k-1), k-1]) have j inside their clamped window: bit (m + k - 1)
__device__ __forceinline__ uint32_t na_inverse_mask(int j, int len, int ksize, int dil) {
    uint32_t mask = 0;
    for (int m = -(ksize - 1); m <= ksize - 1; ++m) {
        const int i = j + m * dil;
        if (i < 0 || i >= len) continue;
        const int s = na_window_start(i, len, ksize, dil);
        if (s <= j && j <= s + (ksize - 1) * dil) mask |= 1u << (m + ksize - 1);
    }
    return mask;
}

// pass B: dk, dv (written into the k and v thirds of dqkv)
template <typename T, int LPH>
__global__ void __launch_bounds__(NA_TILE_THREADS) na2d_bwd_dkv_tile_kernel(const T* __restrict__ qkv, const T* __restrict__ dout,
                                                                           const float* __restrict__ lse, const float* __restrict__ dvec,
                                                                           T* __restrict__ dqkv, NaTile g) {
    CNB_PDL_SYNC();
    constexpr int V = cnb_vec<T>::N;
    CNB_DYN_SMEM(sm_raw);
    T* sm = reinterpret_cast<T*>(sm_raw);
    const int head = blockIdx.y;
    int t = blockIdx.x;
    const int tx = t % g.tiles_x;
    t /= g.tiles_x;
    const int ty = t % g.tiles_y;
    const int b = t / g.tiles_y;
    const int C = g.heads * g.hd;
    const int halo = (g.ksize / 2) * g.dil;
    const int y0 = ty * NA_TH, x0 = tx * NA_TW;
    const int ry0 = na_region_origin(y0, halo, g.H, g.RH), rx0 = na_region_origin(x0, halo, g.W, g.RW);
    const long img_pix0 = (long)b * g.H * g.W;
    // region rows: q_i | dout_i (both hd wide); q and dout have different pixel strides, so stage them one after the other
    {
        const int parts = g.hd / V;
        const int total = g.RH * g.RW * 2 * parts;
        for (int i = threadIdx.x; i < total; i += blockDim.x) {
            const int part = i % (2 * parts);
            const int r = i / (2 * parts);
            const int rx = r % g.RW, ry = r / g.RW;
            const long pix = img_pix0 + (long)(ry0 + ry) * g.W + (rx0 + rx);
            const T* src = part < parts ? qkv + pix * 3 * C + head * g.hd + part * V : dout + pix * C + head * g.hd + (part - parts) * V;
            cnb_cp_async16(sm + (long)r * 2 * g.hd + part * V, src);
        }
    }
    float* sm_l = reinterpret_cast<float*>(sm + (long)g.RH * g.RW * 2 * g.hd);  // lse and D of the region
    float* sm_d = sm_l + g.RH * g.RW;
    for (int r = threadIdx.x; r < g.RH * g.RW; r += blockDim.x) {
        const int rx = r % g.RW, ry = r / g.RW;
        const long pix = img_pix0 + (long)(ry0 + ry) * g.W + (rx0 + rx);
        sm_l[r] = lse[pix * g.heads + head];
        sm_d[r] = dvec[pix * g.heads + head];
    }
    cnb_cp_async_wait_all();
    __syncthreads();

    const int items = NA_TH * NA_TW * LPH;
    for (int it = threadIdx.x; it < items; it += blockDim.x) {
        const int sub = it % LPH;
        const int pl = it / LPH;
        const int lx = pl % NA_TW, ly = pl / NA_TW;
        const bool valid = (y0 + ly) < g.H && (x0 + lx) < g.W;
        const int y = (y0 + ly) < g.H ? y0 + ly : g.H - 1, x = (x0 + lx) < g.W ? x0 + lx : g.W - 1;
        const long pix = img_pix0 + (long)y * g.W + x;
        float kj[V], vj[V], dk[V], dv[V];
        cnb_ldv(qkv + pix * 3 * C + C + head * g.hd + sub * V, kj);
        cnb_ldv(qkv + pix * 3 * C + 2 * C + head * g.hd + sub * V, vj);
#pragma unroll
        for (int j = 0; j < V; ++j) dk[j] = 0.f, dv[j] = 0.f;
        const uint32_t ymask = na_inverse_mask(y, g.H, g.ksize, g.dil), xmask = na_inverse_mask(x, g.W, g.ksize, g.dil);
        for (int my = 0; my < 2 * g.ksize - 1; ++my) {
            if (!((ymask >> my) & 1u)) continue;
            const int iy = y + (my - (g.ksize - 1)) * g.dil;
            for (int mx = 0; mx < 2 * g.ksize - 1; ++mx) {
                if (!((xmask >> mx) & 1u)) continue;
                const int ix = x + (mx - (g.ksize - 1)) * g.dil;
                // a query clamped at the image border can attend a key up to (k-1)*d away, i.e. outside the staged region of an
                // interior-side tile: those few candidates are read from global memory
                const int ry = iy - ry0, rx = ix - rx0;
                const bool staged = ry >= 0 && ry < g.RH && rx >= 0 && rx < g.RW;
                const int r = ry * g.RW + rx;
                const long ipix = img_pix0 + (long)iy * g.W + ix;
                const T* qp = staged ? sm + (long)r * 2 * g.hd + sub * V : qkv + ipix * 3 * C + head * g.hd + sub * V;
                const T* gp = staged ? qp + g.hd : dout + ipix * C + head * g.hd + sub * V;
                const float Li = staged ? sm_l[r] : lse[ipix * g.heads + head];
                const float Di = staged ? sm_d[r] : dvec[ipix * g.heads + head];
                float qi[V], gi[V];
                cnb_ldv(qp, qi);
                cnb_ldv(gp, gi);
                float ps = 0.f, pd = 0.f;
#pragma unroll
                for (int j = 0; j < V; ++j) {
                    ps = fmaf(qi[j], kj[j], ps);
                    pd = fmaf(gi[j], vj[j], pd);
                }
                const float sc = na_group_sum<LPH>(ps) * g.scale, dp = na_group_sum<LPH>(pd);
                const float p = cnb_exp(sc - Li);
                const float ds = p * (dp - Di) * g.scale;
#pragma unroll
                for (int j = 0; j < V; ++j) {
                    dk[j] = fmaf(ds, qi[j], dk[j]);
                    dv[j] = fmaf(p, gi[j], dv[j]);
                }
            }
        }
        if (valid) {
            cnb_stv(dqkv + pix * 3 * C + C + head * g.hd + sub * V, dk);
            cnb_stv(dqkv + pix * 3 * C + 2 * C + head * g.hd + sub * V, dv);
        }
    }
}

}  // namespace cnb
