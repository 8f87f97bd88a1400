// Neighbourhood attention on the tensor cores (bf16, head_dim 32 / 64, kernel 3 / 5 / 7, any dilation).
//
// natten 0.17.1 semantics as configured by the reference (nn/modules/convolution.py:341-350; window rule restated in
// oracle/natten_ref.py): pixel (i, j) attends over the k x k window of its own dilation group, clamped inside the group.  A pixel only
// sees pixels of its residue class (y mod d, x mod d), so dilation d = plain dilation-1 attention on each of the d*d sub-images
// x[gy::d, gx::d]; every kernel below works in SUB-IMAGE coordinates (ys, xs) <-> image pixel (gy + d*ys, gx + d*xs).
//
// Formulation.  The SIMT kernels (k_na_fast.cuh) spend ~27 FMA / shuffle instructions per lane and neighbour and are issue-bound
// (0.05 - 0.4 of the HBM roofline).  Here a warp owns a PATCH of 16 queries = 2 rows x 8 columns of the sub-image and walks the rows of
// the patch's key region (the union of its queries' windows: at most k+1 rows x 16 columns for k <= 9).  For one region row:
//     S[16 q x 16 keys] = Q . K_row^T          2 (n-tiles) x hd/16 mma.m16n8k16
//     P = masked softmax piece (online max / sum, FlashAttention-2 style), packed to bf16 in registers as the next A operand
//     O[16 q x hd]    += P . V_row             hd/8 mma.m16n8k16 (one 16-key k-step)
// so a query costs 128 mma per 16 queries at k = 7 / hd = 64 (38 % of the issued tensor FLOPs are useful) instead of ~1300 FMA + 150
// shuffles per lane.  K / V of the CTA tile + halo are staged ONCE in shared memory (cp.async, 16-byte chunks, XOR swizzle so that
// ldmatrix is conflict-free); Q / dO / O are read once per CTA straight from global memory in the mma fragment layout (a quad reads a
// contiguous 16 bytes, two instructions complete a 32-byte sector).
//
// Backward = two recompute passes, no atomics and no probability records (the SIMT version stored 4 bytes per (pixel, head,
// neighbour): 392 bytes per pixel and head at k = 7):
//   query side (dq):  per region row  S = QK^T, P = exp(S - lse), dP = dO V^T, dS = P (dP - D) scale, dQ += dS K;  also writes
//                     D = rowsum(dO . O) for the other pass
//   key side (dkv):   a warp owns 16 KEYS (2 x 8) and walks the rows of the queries that can see them (clamped windows make that set
//                     up to k + k/2 wide at the borders):  S^T = K Q^T, P^T = exp(S^T - lse_q), dP^T = V dO^T, dS^T = P^T (dP^T - D_q) scale,
//                     dV += P^T dO, dK += dS^T Q.
#pragma once
#include "cnb_common.cuh"

namespace natc {

// ---------------------------------------------------------------------------------------------
// warp-level tensor-core primitives (PTX on the device; the CPU test interpreter emulates them with warp scratch + __syncwarp)
// ---------------------------------------------------------------------------------------------
#ifdef CNB_EMU
struct EmuWarpScratch {
    uint64_t addr[32];
    uint32_t a[32][4], b[32][2];
};
inline EmuWarpScratch& emu_scratch() {
    static thread_local EmuWarpScratch s[32];
    return s[threadIdx.x >> 5];
}
#endif

// four 8x8 b16 matrices; lane i supplies the row address of matrix i/8, row i%8; r[j] = matrix j, element (row lane/4, cols 2*(lane%4), +1)
__device__ __forceinline__ void ldmatrix_x4(uint32_t (&r)[4], const void* row_ptr) {
#ifdef CNB_EMU
    EmuWarpScratch& s = emu_scratch();
    const int lane = threadIdx.x & 31;
    s.addr[lane] = (uint64_t)(uintptr_t)row_ptr;
    __syncwarp();
    for (int j = 0; j < 4; ++j) {
        const unsigned char* p = (const unsigned char*)(uintptr_t)s.addr[8 * j + (lane >> 2)];
        memcpy(&r[j], p + 4 * (lane & 3), 4);
    }
    __syncwarp();
#else
    const uint32_t a = (uint32_t)__cvta_generic_to_shared(row_ptr);
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(a));
#endif
}
// transposed distribution: r[j] = matrix j, elements (row 2*(lane%4), col lane/4) and (row 2*(lane%4)+1, col lane/4)
__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t (&r)[4], const void* row_ptr) {
#ifdef CNB_EMU
    EmuWarpScratch& s = emu_scratch();
    const int lane = threadIdx.x & 31;
    s.addr[lane] = (uint64_t)(uintptr_t)row_ptr;
    __syncwarp();
    for (int j = 0; j < 4; ++j) {
        const unsigned char* p0 = (const unsigned char*)(uintptr_t)s.addr[8 * j + 2 * (lane & 3)];
        const unsigned char* p1 = (const unsigned char*)(uintptr_t)s.addr[8 * j + 2 * (lane & 3) + 1];
        uint16_t lo, hi;
        memcpy(&lo, p0 + 2 * (lane >> 2), 2);
        memcpy(&hi, p1 + 2 * (lane >> 2), 2);
        r[j] = (uint32_t)lo | ((uint32_t)hi << 16);
    }
    __syncwarp();
#else
    const uint32_t a = (uint32_t)__cvta_generic_to_shared(row_ptr);
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
                 : "r"(a));
#endif
}

// D[16x8] += A[16x16] B[16x8], bf16 operands, fp32 accumulate (PTX fragment layouts; g = lane/4, t = lane%4:
// a0 (g, 2t..) a1 (g+8, 2t..) a2 (g, 2t+8..) a3 (g+8, 2t+8..); b0 (k 2t.., n g) b1 (k 2t+8.., n g); c0 c1 (g, 2t..) c2 c3 (g+8, 2t..))
__device__ __forceinline__ void mma16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
#ifdef CNB_EMU
    EmuWarpScratch& s = emu_scratch();
    const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    for (int j = 0; j < 4; ++j) s.a[lane][j] = a[j];
    s.b[lane][0] = b0;
    s.b[lane][1] = b1;
    __syncwarp();
    auto A = [&](int row, int k) {  // element (row, k) of the 16x16 A tile
        const int l = (row & 7) * 4 + ((k & 7) >> 1), reg = (row >> 3) + 2 * (k >> 3);
        const uint32_t w = s.a[l][reg];
        return cnb_bits2f((k & 1) ? (w & 0xffff0000u) : (w << 16));
    };
    auto Bm = [&](int k, int n) {
        const int l = n * 4 + ((k & 7) >> 1), reg = k >> 3;
        const uint32_t w = s.b[l][reg];
        return cnb_bits2f((k & 1) ? (w & 0xffff0000u) : (w << 16));
    };
    float d[4] = {c[0], c[1], c[2], c[3]};
    for (int k = 0; k < 16; ++k) {
        d[0] = fmaf(A(g, k), Bm(k, 2 * t), d[0]);
        d[1] = fmaf(A(g, k), Bm(k, 2 * t + 1), d[1]);
        d[2] = fmaf(A(g + 8, k), Bm(k, 2 * t), d[2]);
        d[3] = fmaf(A(g + 8, k), Bm(k, 2 * t + 1), d[3]);
    }
    __syncwarp();
    for (int j = 0; j < 4; ++j) c[j] = d[j];
#else
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
#endif
}

__device__ __forceinline__ float fast_exp2(float x) {
#ifdef CNB_EMU
    return exp2f(x);
#else
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
#endif
}
__device__ __forceinline__ float quad_max(float v) {
    v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 1));
    return fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 2));
}
__device__ __forceinline__ float quad_sum(float v) {
    v += __shfl_xor_sync(0xffffffffu, v, 1);
    return v + __shfl_xor_sync(0xffffffffu, v, 2);
}
__device__ __forceinline__ int imin(int a, int b) { return a < b ? a : b; }
__device__ __forceinline__ uint32_t ldg32(const bf16_t* p) { return *reinterpret_cast<const uint32_t*>(p); }

// ---------------------------------------------------------------------------------------------
// geometry
// ---------------------------------------------------------------------------------------------
struct Geom {
    int B, H, W, heads, dil;
    int tiles_y, tiles_x;  // tiles per sub-image (sized for the largest sub-image)
    float scale;
};
constexpr int TW = 16;       // tile width (sub-image columns): two 8-column patches
constexpr int PATCH_W = 8;   // a patch = 2 rows x 8 columns = the 16 rows of an mma tile

// first window coordinate of query p on an axis of length L (natten's clamped window)
template <int K>
__device__ __forceinline__ int win_start(int p, int L) {
    const int s = p - K / 2;
    return s < 0 ? 0 : (s > L - K ? L - K : s);
}
// first / last query whose window contains key k (contiguous because win_start is monotone)
template <int K>
__device__ __forceinline__ int q_lo(int k) {
    return k <= K - 1 ? 0 : k - K / 2;
}
template <int K>
__device__ __forceinline__ int q_hi(int k, int L) {
    return k >= L - K ? L - 1 : k + K / 2;
}

// shared-memory pixel rows of HD bf16 (HD*2 bytes) with the 16-byte chunk index XOR-swizzled by the pixel index: 8 consecutive pixels
// read at one logical chunk (an ldmatrix 8x8 matrix) hit 8 different bank groups
template <int HD>
__device__ __forceinline__ int swz(int pix, int chunk) {
    return HD == 64 ? (pix * 8 + (chunk ^ (pix & 7))) : (pix * 4 + (chunk ^ ((pix >> 1) & 3)));  // in 16-byte units
}

// Stage the HD channels at `chan` of the sub-image pixels [y0, y0+nrows) x [x0, x0+ncols) into `dst` (pixel pitch `pitch`; pixels
// outside the range or the sub-image become zero).  All threads of the CTA call it; completion = cp_async_wait + __syncthreads.
template <int HD, int NT>
__device__ __forceinline__ void stage_region(uint4* dst, const bf16_t* __restrict__ base, long pix_stride, int chan, int gy, int gx, int dil,
                                             int Wimg, int y0, int x0, int nrows, int ncols, int pitch, int alloc_pix) {
    constexpr int CH = HD / 8;
    for (int i = threadIdx.x; i < alloc_pix * CH; i += NT) {
        const int pix = i / CH, c = i - pix * CH;
        const int r = pix / pitch, col = pix - r * pitch;
        uint4* d = dst + swz<HD>(pix, c);
        if (r < nrows && col < ncols) {
            const long gp = (long)(gy + dil * (y0 + r)) * Wimg + (gx + dil * (x0 + col));
            cnb_cp_async16(d, base + gp * pix_stride + chan + c * 8);
        } else {
            *d = make_uint4(0u, 0u, 0u, 0u);
        }
    }
}

// A fragments (16 rows x HD) of a patch straight from global memory: row g -> pixel p0, row g+8 -> pixel p1 (nullptr = zeros)
template <int HD>
__device__ __forceinline__ void load_a_frags(uint32_t (&a)[HD / 16][4], const bf16_t* p0, const bf16_t* p1, int t) {
#pragma unroll
    for (int ks = 0; ks < HD / 16; ++ks) {
        a[ks][0] = p0 ? ldg32(p0 + ks * 16 + 2 * t) : 0u;
        a[ks][1] = p1 ? ldg32(p1 + ks * 16 + 2 * t) : 0u;
        a[ks][2] = p0 ? ldg32(p0 + ks * 16 + 8 + 2 * t) : 0u;
        a[ks][3] = p1 ? ldg32(p1 + ks * 16 + 8 + 2 * t) : 0u;
    }
}

// S[2 n-tiles][4] = A(16 x HD) . rows^T for the 16 consecutive shared-memory pixels starting at `pix0` (B operand, non-transposed)
template <int HD>
__device__ __forceinline__ void mma_qk(float (&s)[2][4], const uint32_t (&a)[HD / 16][4], const uint4* sm, int pix0, int lane) {
#pragma unroll
    for (int nt = 0; nt < 2; ++nt)
#pragma unroll
        for (int j = 0; j < 4; ++j) s[nt][j] = 0.f;
    const int key = pix0 + (lane & 7) + 8 * (lane >> 4), half = (lane >> 3) & 1;
#pragma unroll
    for (int ks = 0; ks < HD / 16; ++ks) {
        uint32_t b[4];
        ldmatrix_x4(b, sm + swz<HD>(key, ks * 2 + half));
        mma16816(s[0], a[ks], b[0], b[1]);
        mma16816(s[1], a[ks], b[2], b[3]);
    }
}

// acc[HD/8][4] += P(16 x 16, A fragment) . rows for the 16 consecutive shared-memory pixels starting at `pix0` (B operand, transposed)
template <int HD>
__device__ __forceinline__ void mma_pv(float (&acc)[HD / 8][4], const uint32_t (&p)[4], const uint4* sm, int pix0, int lane) {
    const int key = pix0 + (lane & 7) + 8 * ((lane >> 3) & 1), half = lane >> 4;
#pragma unroll
    for (int hn = 0; hn < HD / 16; ++hn) {
        uint32_t b[4];
        ldmatrix_x4_trans(b, sm + swz<HD>(key, hn * 2 + half));
        mma16816(acc[2 * hn], p, b[0], b[1]);
        mma16816(acc[2 * hn + 1], p, b[2], b[3]);
    }
}

// ---------------------------------------------------------------------------------------------
// tile decoding shared by the three kernels: blockIdx.x = (dilation group, tile), y = head, z = batch sample
// ---------------------------------------------------------------------------------------------
struct TileCtx {
    int b, h, gy, gx, Hs, Ws, ty0, tx0;
    bool empty;
};
template <int TH>
__device__ __forceinline__ TileCtx decode_tile(const Geom& g) {
    TileCtx c;
    const int tiles = g.tiles_y * g.tiles_x;
    const int grp = blockIdx.x / tiles, tile = blockIdx.x - grp * tiles;
    c.gy = grp / g.dil, c.gx = grp - c.gy * g.dil;
    c.Hs = (g.H - c.gy + g.dil - 1) / g.dil, c.Ws = (g.W - c.gx + g.dil - 1) / g.dil;
    const int tyi = tile / g.tiles_x, txi = tile - tyi * g.tiles_x;
    c.ty0 = tyi * TH, c.tx0 = txi * TW;
    c.h = blockIdx.y, c.b = blockIdx.z;
    c.empty = c.ty0 >= c.Hs || c.tx0 >= c.Ws;
    return c;
}

// ---------------------------------------------------------------------------------------------
// forward.  CTA = TH x 16 queries of one sub-image, head and sample; TH/2 * 2 warps, one patch each.
// shared memory: K and V of the tile's key region, (TH + K - 1) rows x 24 columns
// ---------------------------------------------------------------------------------------------
template <int K, int TH>
struct FwdCfg {
    static constexpr int RH = TH + K - 1, RWP = TW + 8;  // a patch reads 16 columns starting at most 8 columns into the region
    static constexpr int PIX = RH * RWP;
    static constexpr int NWARPS = TH;  // (TH/2 patch rows) x 2 patch columns
    static constexpr int NT = NWARPS * 32;
    static_assert(K <= 9 && (K & 1), "window sizes up to 9");
    static constexpr size_t smem(int hd) { return (size_t)2 * PIX * hd * 2; }
};

template <int HD, int K, int TH>
__global__ void __launch_bounds__(FwdCfg<K, TH>::NT) na_tc_fwd_kernel(const bf16_t* __restrict__ qkv, bf16_t* __restrict__ out,
                                                                     float* __restrict__ lse, Geom g) {
    using Cfg = FwdCfg<K, TH>;
    CNB_DYN_SMEM(smem_raw);
    uint4* sK = reinterpret_cast<uint4*>(smem_raw);
    uint4* sV = sK + Cfg::PIX * (HD / 8);
    const TileCtx c = decode_tile<TH>(g);
    if (c.empty) return;
    CNB_PDL_SYNC();
    const int Cn = g.heads * HD;
    const long pstride = 3L * Cn;
    const bf16_t* img = qkv + (long)c.b * g.H * g.W * pstride;
    const int ty1 = imin(c.ty0 + TH, c.Hs) - 1, tx1 = imin(c.tx0 + TW, c.Ws) - 1;
    const int ry0 = win_start<K>(c.ty0, c.Hs), ry1 = win_start<K>(ty1, c.Hs) + K - 1;
    const int rx0 = win_start<K>(c.tx0, c.Ws), rx1 = win_start<K>(tx1, c.Ws) + K - 1;
    stage_region<HD, Cfg::NT>(sK, img, pstride, Cn + c.h * HD, c.gy, c.gx, g.dil, g.W, ry0, rx0, ry1 - ry0 + 1, rx1 - rx0 + 1, Cfg::RWP, Cfg::PIX);
    stage_region<HD, Cfg::NT>(sV, img, pstride, 2 * Cn + c.h * HD, c.gy, c.gx, g.dil, g.W, ry0, rx0, ry1 - ry0 + 1, rx1 - rx0 + 1, Cfg::RWP, Cfg::PIX);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, gq = lane >> 2, t = lane & 3;
    const int py = c.ty0 + (warp >> 1) * 2, px = c.tx0 + (warp & 1) * PATCH_W;
    const bool patch_live = py < c.Hs && px < c.Ws;
    // this lane's two queries: rows gq (patch row 0) and gq + 8 (patch row 1), column px + gq
    const int qx = px + gq;
    const int qy0 = py, qy1 = py + 1;
    const bool v0 = patch_live && qx < c.Ws, v1 = v0 && qy1 < c.Hs;
    auto pix_ptr = [&](int ys, int xs) { return img + ((long)(c.gy + g.dil * ys) * g.W + (c.gx + g.dil * xs)) * pstride; };
    uint32_t qa[HD / 16][4];
    load_a_frags<HD>(qa, v0 ? pix_ptr(qy0, qx) + c.h * HD : nullptr, v1 ? pix_ptr(qy1, qx) + c.h * HD : nullptr, t);
    cnb_cp_async_wait_all();
    __syncthreads();
    if (!patch_live) return;

    const int wy0 = win_start<K>(qy0, c.Hs), wy1 = win_start<K>(imin(qy1, c.Hs - 1), c.Hs);
    const int wx = win_start<K>(imin(qx, c.Ws - 1), c.Ws);
    const int cx0 = win_start<K>(px, c.Ws);                   // first key column the patch reads (16 columns from here)
    const int r_first = wy0, r_last = wy1 + K - 1;            // key rows of the patch (uniform over the warp)
    bool okx[2][2];
#pragma unroll
    for (int nt = 0; nt < 2; ++nt)
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const int kx = cx0 + nt * 8 + 2 * t + j;
            okx[nt][j] = kx >= wx && kx < wx + K;
        }
    const float sc2 = g.scale * 1.4426950408889634f;  // logits in log2 units
    float m_run[2] = {-INFINITY, -INFINITY}, l_run[2] = {0.f, 0.f};
    float o[HD / 8][4];
#pragma unroll
    for (int n = 0; n < HD / 8; ++n)
#pragma unroll
        for (int j = 0; j < 4; ++j) o[n][j] = 0.f;

    for (int ky = r_first; ky <= r_last; ++ky) {
        const int pix0 = (ky - ry0) * Cfg::RWP + (cx0 - rx0);
        float s[2][4];
        mma_qk<HD>(s, qa, sK, pix0, lane);
        const bool oky0 = ky >= wy0 && ky < wy0 + K, oky1 = ky >= wy1 && ky < wy1 + K;
        float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
        for (int nt = 0; nt < 2; ++nt)
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                s[nt][j] = (oky0 && okx[nt][j]) ? s[nt][j] * sc2 : -INFINITY;
                s[nt][2 + j] = (oky1 && okx[nt][j]) ? s[nt][2 + j] * sc2 : -INFINITY;
                mx0 = fmaxf(mx0, s[nt][j]);
                mx1 = fmaxf(mx1, s[nt][2 + j]);
            }
        mx0 = quad_max(mx0), mx1 = quad_max(mx1);
        const float mn0 = fmaxf(m_run[0], mx0), mn1 = fmaxf(m_run[1], mx1);
        const float ms0 = mn0 == -INFINITY ? 0.f : mn0, ms1 = mn1 == -INFINITY ? 0.f : mn1;  // nothing seen yet: keep exp2(-inf) = 0
        const float al0 = fast_exp2(m_run[0] - ms0), al1 = fast_exp2(m_run[1] - ms1);
        m_run[0] = mn0, m_run[1] = mn1;
        float p[2][4];
        float ps0 = 0.f, ps1 = 0.f;
#pragma unroll
        for (int nt = 0; nt < 2; ++nt)
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                p[nt][j] = fast_exp2(s[nt][j] - ms0);
                p[nt][2 + j] = fast_exp2(s[nt][2 + j] - ms1);
                ps0 += p[nt][j];
                ps1 += p[nt][2 + j];
            }
        l_run[0] = l_run[0] * al0 + ps0;
        l_run[1] = l_run[1] * al1 + ps1;
#pragma unroll
        for (int n = 0; n < HD / 8; ++n) {
            o[n][0] *= al0, o[n][1] *= al0;
            o[n][2] *= al1, o[n][3] *= al1;
        }
        const uint32_t pa[4] = {cnb_pack_bf16x2(p[0][0], p[0][1]), cnb_pack_bf16x2(p[0][2], p[0][3]), cnb_pack_bf16x2(p[1][0], p[1][1]),
                                cnb_pack_bf16x2(p[1][2], p[1][3])};
        mma_pv<HD>(o, pa, sV, pix0, lane);
    }
    const float l0 = quad_sum(l_run[0]), l1 = quad_sum(l_run[1]);
    const float i0 = 1.f / l0, i1 = 1.f / l1;
    if (v0) {
        bf16_t* op = out + (((long)c.b * g.H + (c.gy + g.dil * qy0)) * g.W + (c.gx + g.dil * qx)) * Cn + c.h * HD;
#pragma unroll
        for (int n = 0; n < HD / 8; ++n) *reinterpret_cast<uint32_t*>(op + n * 8 + 2 * t) = cnb_pack_bf16x2(o[n][0] * i0, o[n][1] * i0);
        if (t == 0)
            lse[(((long)c.b * g.H + (c.gy + g.dil * qy0)) * g.W + (c.gx + g.dil * qx)) * g.heads + c.h] = (m_run[0] + log2f(l0)) * 0.6931471805599453f;
    }
    if (v1) {
        bf16_t* op = out + (((long)c.b * g.H + (c.gy + g.dil * qy1)) * g.W + (c.gx + g.dil * qx)) * Cn + c.h * HD;
#pragma unroll
        for (int n = 0; n < HD / 8; ++n) *reinterpret_cast<uint32_t*>(op + n * 8 + 2 * t) = cnb_pack_bf16x2(o[n][2] * i1, o[n][3] * i1);
        if (t == 0)
            lse[(((long)c.b * g.H + (c.gy + g.dil * qy1)) * g.W + (c.gx + g.dil * qx)) * g.heads + c.h] = (m_run[1] + log2f(l1)) * 0.6931471805599453f;
    }
}

// ---------------------------------------------------------------------------------------------
// backward, query side: dq[16 q x HD] and D = rowsum(dO . O).  Same tiling and staging as the forward.
// ---------------------------------------------------------------------------------------------
template <int HD, int K, int TH>
__global__ void __launch_bounds__(FwdCfg<K, TH>::NT) na_tc_bwd_dq_kernel(const bf16_t* __restrict__ qkv, const bf16_t* __restrict__ dout,
                                                                        const bf16_t* __restrict__ outp, const float* __restrict__ lse,
                                                                        float* __restrict__ dvec, bf16_t* __restrict__ dqkv, Geom g) {
    using Cfg = FwdCfg<K, TH>;
    CNB_DYN_SMEM(smem_raw);
    uint4* sK = reinterpret_cast<uint4*>(smem_raw);
    uint4* sV = sK + Cfg::PIX * (HD / 8);
    const TileCtx c = decode_tile<TH>(g);
    if (c.empty) return;
    CNB_PDL_SYNC();
    const int Cn = g.heads * HD;
    const long pstride = 3L * Cn;
    const bf16_t* img = qkv + (long)c.b * g.H * g.W * pstride;
    const int ty1 = imin(c.ty0 + TH, c.Hs) - 1, tx1 = imin(c.tx0 + TW, c.Ws) - 1;
    const int ry0 = win_start<K>(c.ty0, c.Hs), ry1 = win_start<K>(ty1, c.Hs) + K - 1;
    const int rx0 = win_start<K>(c.tx0, c.Ws), rx1 = win_start<K>(tx1, c.Ws) + K - 1;
    stage_region<HD, Cfg::NT>(sK, img, pstride, Cn + c.h * HD, c.gy, c.gx, g.dil, g.W, ry0, rx0, ry1 - ry0 + 1, rx1 - rx0 + 1, Cfg::RWP, Cfg::PIX);
    stage_region<HD, Cfg::NT>(sV, img, pstride, 2 * Cn + c.h * HD, c.gy, c.gx, g.dil, g.W, ry0, rx0, ry1 - ry0 + 1, rx1 - rx0 + 1, Cfg::RWP, Cfg::PIX);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, gq = lane >> 2, t = lane & 3;
    const int py = c.ty0 + (warp >> 1) * 2, px = c.tx0 + (warp & 1) * PATCH_W;
    const bool patch_live = py < c.Hs && px < c.Ws;
    const int qx = px + gq, qy0 = py, qy1 = py + 1;
    const bool v0 = patch_live && qx < c.Ws, v1 = v0 && qy1 < c.Hs;
    const long gp0 = ((long)c.b * g.H + (c.gy + g.dil * qy0)) * g.W + (c.gx + g.dil * qx);
    const long gp1 = ((long)c.b * g.H + (c.gy + g.dil * qy1)) * g.W + (c.gx + g.dil * qx);
    uint32_t qa[HD / 16][4], da[HD / 16][4];
    load_a_frags<HD>(qa, v0 ? qkv + gp0 * pstride + c.h * HD : nullptr, v1 ? qkv + gp1 * pstride + c.h * HD : nullptr, t);
    load_a_frags<HD>(da, v0 ? dout + gp0 * Cn + c.h * HD : nullptr, v1 ? dout + gp1 * Cn + c.h * HD : nullptr, t);
    float D0 = 0.f, D1 = 0.f;
    {
        uint32_t oa[HD / 16][4];
        load_a_frags<HD>(oa, v0 ? outp + gp0 * Cn + c.h * HD : nullptr, v1 ? outp + gp1 * Cn + c.h * HD : nullptr, t);
#pragma unroll
        for (int ks = 0; ks < HD / 16; ++ks) {
            D0 = cnb_fma2_bf16(da[ks][0], oa[ks][0], D0);
            D0 = cnb_fma2_bf16(da[ks][2], oa[ks][2], D0);
            D1 = cnb_fma2_bf16(da[ks][1], oa[ks][1], D1);
            D1 = cnb_fma2_bf16(da[ks][3], oa[ks][3], D1);
        }
        D0 = quad_sum(D0), D1 = quad_sum(D1);
    }
    const float lse0 = v0 ? lse[gp0 * g.heads + c.h] * 1.4426950408889634f : 0.f;
    const float lse1 = v1 ? lse[gp1 * g.heads + c.h] * 1.4426950408889634f : 0.f;
    if (t == 0) {
        if (v0) dvec[gp0 * g.heads + c.h] = D0;
        if (v1) dvec[gp1 * g.heads + c.h] = D1;
    }
    cnb_cp_async_wait_all();
    __syncthreads();
    if (!patch_live) return;

    const int wy0 = win_start<K>(qy0, c.Hs), wy1 = win_start<K>(imin(qy1, c.Hs - 1), c.Hs);
    const int wx = win_start<K>(imin(qx, c.Ws - 1), c.Ws);
    const int cx0 = win_start<K>(px, c.Ws);
    const int r_first = wy0, r_last = wy1 + K - 1;
    bool okx[2][2];
#pragma unroll
    for (int nt = 0; nt < 2; ++nt)
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const int kx = cx0 + nt * 8 + 2 * t + j;
            okx[nt][j] = kx >= wx && kx < wx + K;
        }
    const float sc2 = g.scale * 1.4426950408889634f;
    float dq[HD / 8][4];
#pragma unroll
    for (int n = 0; n < HD / 8; ++n)
#pragma unroll
        for (int j = 0; j < 4; ++j) dq[n][j] = 0.f;

    for (int ky = r_first; ky <= r_last; ++ky) {
        const int pix0 = (ky - ry0) * Cfg::RWP + (cx0 - rx0);
        float s[2][4], dp[2][4];
        mma_qk<HD>(s, qa, sK, pix0, lane);
        mma_qk<HD>(dp, da, sV, pix0, lane);
        const bool oky0 = v0 && ky >= wy0 && ky < wy0 + K, oky1 = v1 && ky >= wy1 && ky < wy1 + K;
        float ds[2][4];
#pragma unroll
        for (int nt = 0; nt < 2; ++nt)
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                const float p0 = (oky0 && okx[nt][j]) ? fast_exp2(s[nt][j] * sc2 - lse0) : 0.f;
                const float p1 = (oky1 && okx[nt][j]) ? fast_exp2(s[nt][2 + j] * sc2 - lse1) : 0.f;
                ds[nt][j] = p0 * (dp[nt][j] - D0) * g.scale;
                ds[nt][2 + j] = p1 * (dp[nt][2 + j] - D1) * g.scale;
            }
        const uint32_t sa[4] = {cnb_pack_bf16x2(ds[0][0], ds[0][1]), cnb_pack_bf16x2(ds[0][2], ds[0][3]), cnb_pack_bf16x2(ds[1][0], ds[1][1]),
                                cnb_pack_bf16x2(ds[1][2], ds[1][3])};
        mma_pv<HD>(dq, sa, sK, pix0, lane);
    }
    if (v0) {
        bf16_t* dp0 = dqkv + gp0 * pstride + c.h * HD;
#pragma unroll
        for (int n = 0; n < HD / 8; ++n) *reinterpret_cast<uint32_t*>(dp0 + n * 8 + 2 * t) = cnb_pack_bf16x2(dq[n][0], dq[n][1]);
    }
    if (v1) {
        bf16_t* dp1 = dqkv + gp1 * pstride + c.h * HD;
#pragma unroll
        for (int n = 0; n < HD / 8; ++n) *reinterpret_cast<uint32_t*>(dp1 + n * 8 + 2 * t) = cnb_pack_bf16x2(dq[n][2], dq[n][3]);
    }
}

// ---------------------------------------------------------------------------------------------
// backward, key side: dk, dv of TH x 16 keys.  Shared memory: Q and dO of the query region (pixel pitch = the region's own width; a
// 16-column read that runs past a region row reads the next row's finite values and is masked), lse (log2 units) and D per region query.
// ---------------------------------------------------------------------------------------------
template <int K, int TH>
struct DkvCfg {
    static constexpr int RH = TH + K - 1 + K / 2, RW = TW + K - 1 + K / 2;
    static constexpr int PIX = RH * RW + 2 * TW;  // slack for the reads that run past the last row
    static constexpr int NWARPS = TH;
    static constexpr int NT = NWARPS * 32;
    static constexpr size_t smem(int hd) { return (size_t)2 * PIX * hd * 2 + (size_t)2 * PIX * 4; }
};

template <int HD, int K, int TH>
__global__ void __launch_bounds__(DkvCfg<K, TH>::NT) na_tc_bwd_dkv_kernel(const bf16_t* __restrict__ qkv, const bf16_t* __restrict__ dout,
                                                                         const float* __restrict__ lse, const float* __restrict__ dvec,
                                                                         bf16_t* __restrict__ dqkv, Geom g) {
    using Cfg = DkvCfg<K, TH>;
    CNB_DYN_SMEM(smem_raw);
    uint4* sQ = reinterpret_cast<uint4*>(smem_raw);
    uint4* sO = sQ + Cfg::PIX * (HD / 8);
    float* sL = reinterpret_cast<float*>(sO + Cfg::PIX * (HD / 8));
    float* sD = sL + Cfg::PIX;
    const TileCtx c = decode_tile<TH>(g);
    if (c.empty) return;
    CNB_PDL_SYNC();
    const int Cn = g.heads * HD;
    const long pstride = 3L * Cn;
    const long img_pix0 = (long)c.b * g.H * g.W;
    const int ty1 = imin(c.ty0 + TH, c.Hs) - 1, tx1 = imin(c.tx0 + TW, c.Ws) - 1;
    const int ry0 = q_lo<K>(c.ty0), ry1 = q_hi<K>(ty1, c.Hs);
    const int rx0 = q_lo<K>(c.tx0), rx1 = q_hi<K>(tx1, c.Ws);
    const int nrows = ry1 - ry0 + 1, RW = rx1 - rx0 + 1;  // RW <= Cfg::RW
    stage_region<HD, Cfg::NT>(sQ, qkv + img_pix0 * pstride, pstride, c.h * HD, c.gy, c.gx, g.dil, g.W, ry0, rx0, nrows, RW, RW, Cfg::PIX);
    stage_region<HD, Cfg::NT>(sO, dout + img_pix0 * Cn, (long)Cn, c.h * HD, c.gy, c.gx, g.dil, g.W, ry0, rx0, nrows, RW, RW, Cfg::PIX);
    for (int i = threadIdx.x; i < Cfg::PIX; i += Cfg::NT) {
        const int r = i / RW, col = i - r * RW;
        float l = 0.f, d = 0.f;
        if (r < nrows) {
            const long gp = img_pix0 + (long)(c.gy + g.dil * (ry0 + r)) * g.W + (c.gx + g.dil * (rx0 + col));
            l = lse[gp * g.heads + c.h] * 1.4426950408889634f;
            d = dvec[gp * g.heads + c.h];
        }
        sL[i] = l, sD[i] = d;
    }

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, gk = lane >> 2, t = lane & 3;
    const int py = c.ty0 + (warp >> 1) * 2, px = c.tx0 + (warp & 1) * PATCH_W;
    const bool patch_live = py < c.Hs && px < c.Ws;
    const int kx = px + gk, ky0 = py, ky1 = py + 1;  // this lane's two keys: mma rows gk and gk + 8
    const bool v0 = patch_live && kx < c.Ws, v1 = v0 && ky1 < c.Hs;
    const long gp0 = img_pix0 + (long)(c.gy + g.dil * ky0) * g.W + (c.gx + g.dil * kx);
    const long gp1 = img_pix0 + (long)(c.gy + g.dil * ky1) * g.W + (c.gx + g.dil * kx);
    uint32_t ka[HD / 16][4], va[HD / 16][4];
    load_a_frags<HD>(ka, v0 ? qkv + gp0 * pstride + Cn + c.h * HD : nullptr, v1 ? qkv + gp1 * pstride + Cn + c.h * HD : nullptr, t);
    load_a_frags<HD>(va, v0 ? qkv + gp0 * pstride + 2 * Cn + c.h * HD : nullptr, v1 ? qkv + gp1 * pstride + 2 * Cn + c.h * HD : nullptr, t);
    cnb_cp_async_wait_all();
    __syncthreads();
    if (!patch_live) return;

    // queries that can see the patch's keys: rows [qr0, qr1], columns [qc0, qc1] (uniform over the warp)
    const int pyl = imin(py + 1, c.Hs - 1), pxl = imin(px + PATCH_W - 1, c.Ws - 1);
    const int qr0 = q_lo<K>(py), qr1 = q_hi<K>(pyl, c.Hs);
    const int qc0 = q_lo<K>(px), qc1 = q_hi<K>(pxl, c.Ws);
    const float sc2 = g.scale * 1.4426950408889634f;
    float dk[HD / 8][4], dv[HD / 8][4];
#pragma unroll
    for (int n = 0; n < HD / 8; ++n)
#pragma unroll
        for (int j = 0; j < 4; ++j) dk[n][j] = 0.f, dv[n][j] = 0.f;

    for (int qy = qr0; qy <= qr1; ++qy) {
        const int wy = win_start<K>(qy, c.Hs);
        const bool oky0 = v0 && ky0 >= wy && ky0 < wy + K, oky1 = v1 && ky1 >= wy && ky1 < wy + K;
        for (int qcb = qc0; qcb <= qc1; qcb += 16) {  // one or (at the clamped borders) two 16-query column chunks
            const int pix0 = (qy - ry0) * RW + (qcb - rx0);
            float s[2][4], dp[2][4];
            mma_qk<HD>(s, ka, sQ, pix0, lane);
            mma_qk<HD>(dp, va, sO, pix0, lane);
            float pt[2][4], ds[2][4];
#pragma unroll
            for (int nt = 0; nt < 2; ++nt)
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    const int col = nt * 8 + 2 * t + j, qx = qcb + col;
                    const int wxq = win_start<K>(imin(qx, c.Ws - 1), c.Ws);
                    const bool okq = qx <= qc1 && kx >= wxq && kx < wxq + K;
                    const float lq = sL[pix0 + col], dq = sD[pix0 + col];
                    const float p0 = (oky0 && okq) ? fast_exp2(s[nt][j] * sc2 - lq) : 0.f;
                    const float p1 = (oky1 && okq) ? fast_exp2(s[nt][2 + j] * sc2 - lq) : 0.f;
                    pt[nt][j] = p0, pt[nt][2 + j] = p1;
                    ds[nt][j] = p0 * (dp[nt][j] - dq) * g.scale;
                    ds[nt][2 + j] = p1 * (dp[nt][2 + j] - dq) * g.scale;
                }
            const uint32_t pa[4] = {cnb_pack_bf16x2(pt[0][0], pt[0][1]), cnb_pack_bf16x2(pt[0][2], pt[0][3]), cnb_pack_bf16x2(pt[1][0], pt[1][1]),
                                    cnb_pack_bf16x2(pt[1][2], pt[1][3])};
            const uint32_t sa[4] = {cnb_pack_bf16x2(ds[0][0], ds[0][1]), cnb_pack_bf16x2(ds[0][2], ds[0][3]), cnb_pack_bf16x2(ds[1][0], ds[1][1]),
                                    cnb_pack_bf16x2(ds[1][2], ds[1][3])};
            mma_pv<HD>(dv, pa, sO, pix0, lane);
            mma_pv<HD>(dk, sa, sQ, pix0, lane);
        }
    }
    if (v0) {
        bf16_t* k0 = dqkv + gp0 * pstride + Cn + c.h * HD;
#pragma unroll
        for (int n = 0; n < HD / 8; ++n) {
            *reinterpret_cast<uint32_t*>(k0 + n * 8 + 2 * t) = cnb_pack_bf16x2(dk[n][0], dk[n][1]);
            *reinterpret_cast<uint32_t*>(k0 + Cn + n * 8 + 2 * t) = cnb_pack_bf16x2(dv[n][0], dv[n][1]);
        }
    }
    if (v1) {
        bf16_t* k1 = dqkv + gp1 * pstride + Cn + c.h * HD;
#pragma unroll
        for (int n = 0; n < HD / 8; ++n) {
            *reinterpret_cast<uint32_t*>(k1 + n * 8 + 2 * t) = cnb_pack_bf16x2(dk[n][2], dk[n][3]);
            *reinterpret_cast<uint32_t*>(k1 + Cn + n * 8 + 2 * t) = cnb_pack_bf16x2(dv[n][2], dv[n][3]);
        }
    }
}

inline bool eligible(int hd, int ksize, int dtype) { return dtype == CNB_BF16 && (hd == 32 || hd == 64) && (ksize == 3 || ksize == 5 || ksize == 7); }

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
#ifdef CNB_EMU
#define NATC_SET_SMEM(kfn, bytes) ((void)0)
#else
#define NATC_SET_SMEM(kfn, bytes) cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(bytes))
#endif
// ---- tensor-core neighbourhood attention (k_na_tc.cuh): bf16, head_dim 32 / 64, kernel 3 / 5 / 7, any dilation ----
// CNB_NA_TC=0 falls back to the SIMT kernels (A/B timing)
inline bool enabled() {
    static const bool on = [] {
        const char* e = getenv("CNB_NA_TC");
        return !(e && e[0] == '0');
    }();
    return on;
}
inline bool usable(int B, int heads, int hd, int ksize, int dtype) {
    // kernel 3: a query has 9 neighbours and the mma formulation's fixed cost per patch (staging, 16-column key rows of which 10 are
    // used) is not paid back -- measured on B200 at BASELINE config 2 (profiles/r2f_bench_na_*): forward 0.40 vs 0.36 ms, backward
    // 1.49 vs 1.08 ms for the SIMT kernels of k_na_fast.cuh, which keep that shape.  CNB_NA_TC=3 forces the tensor-core path there.
    static const int min_k = [] {
        const char* e = getenv("CNB_NA_TC");
        return (e && e[0] == '3') ? 3 : 5;
    }();
    return enabled() && eligible(hd, ksize, dtype) && ksize >= min_k && B <= 65535 && heads <= 65535;
}
template <int TH>
inline Geom make_geom(int B, int H, int W, int heads, int dil, float scale, dim3* grid) {
    Geom g;
    g.B = B, g.H = H, g.W = W, g.heads = heads, g.dil = dil, g.scale = scale;
    const int hs = cnb_div_up(H, dil), ws = cnb_div_up(W, dil);
    g.tiles_y = cnb_div_up(hs, TH), g.tiles_x = cnb_div_up(ws, TW);
    *grid = dim3((unsigned)(dil * dil * g.tiles_y * g.tiles_x), (unsigned)heads, (unsigned)B);
    return g;
}
#define CNB_NATC_K(KS, ...)        \
    do {                            \
        if ((KS) == 3) {            \
            constexpr int KV = 3;   \
            __VA_ARGS__;                 \
        } else if ((KS) == 5) {     \
            constexpr int KV = 5;   \
            __VA_ARGS__;                 \
        } else {                    \
            constexpr int KV = 7;   \
            __VA_ARGS__;                 \
        }                           \
    } while (0)
#define CNB_NATC_HD(HDV, ...)      \
    do {                            \
        if ((HDV) == 64) {          \
            constexpr int HV = 64;  \
            __VA_ARGS__;                 \
        } else {                    \
            constexpr int HV = 32;  \
            __VA_ARGS__;                 \
        }                           \
    } while (0)

inline int launch_fwd(const void* qkv, void* out, float* lse, int B, int H, int W, int heads, int hd, int ksize, int dilation, float scale,
                     void* stream) {
    constexpr int TH = 8;
    dim3 grid;
    const Geom g = make_geom<TH>(B, H, W, heads, dilation, scale, &grid);
    CNB_NATC_K(ksize, CNB_NATC_HD(hd, {
        using Cfg = FwdCfg<KV, TH>;
        const size_t smem = Cfg::smem(HV);
        NATC_SET_SMEM((na_tc_fwd_kernel<HV, KV, TH>), smem);
        CNB_LAUNCH((na_tc_fwd_kernel<HV, KV, TH>), grid, dim3(Cfg::NT), smem, (cudaStream_t)stream, (const bf16_t*)qkv, (bf16_t*)out, lse, g);
    }));
    CNB_CHECK_LAUNCH("na_tc_fwd_kernel");
    return CNB_OK;
}

inline int launch_bwd(const void* qkv, const void* dout, const void* out, const float* lse, float* dvec, void* dqkv, int B, int H, int W,
                     int heads, int hd, int ksize, int dilation, float scale, void* stream) {
    {
        constexpr int TH = 8;
        dim3 grid;
        const Geom g = make_geom<TH>(B, H, W, heads, dilation, scale, &grid);
        CNB_NATC_K(ksize, CNB_NATC_HD(hd, {
            using Cfg = FwdCfg<KV, TH>;
            const size_t smem = Cfg::smem(HV);
            NATC_SET_SMEM((na_tc_bwd_dq_kernel<HV, KV, TH>), smem);
            CNB_LAUNCH((na_tc_bwd_dq_kernel<HV, KV, TH>), grid, dim3(Cfg::NT), smem, (cudaStream_t)stream, (const bf16_t*)qkv,
                       (const bf16_t*)dout, (const bf16_t*)out, lse, dvec, (bf16_t*)dqkv, g);
        }));
    }
    // key side: 16-row tiles (16 warps) where one 8-row CTA per SM would be all that fits (k 7, head_dim 64), 8-row tiles otherwise
    static const int dkv_th = [] {  // CNB_NA_DKV_TH=8 forces the 8-row key tiles everywhere (A/B timing)
        const char* e = getenv("CNB_NA_DKV_TH");
        return (e && e[0] == '8') ? 8 : 16;
    }();
    if (ksize == 7 && hd == 64 && dkv_th == 16) {
        constexpr int TH = 16;
        dim3 grid;
        const Geom g = make_geom<TH>(B, H, W, heads, dilation, scale, &grid);
        using Cfg = DkvCfg<7, TH>;
        const size_t smem = Cfg::smem(64);
        NATC_SET_SMEM((na_tc_bwd_dkv_kernel<64, 7, TH>), smem);
        CNB_LAUNCH((na_tc_bwd_dkv_kernel<64, 7, TH>), grid, dim3(Cfg::NT), smem, (cudaStream_t)stream, (const bf16_t*)qkv, (const bf16_t*)dout,
                   lse, (const float*)dvec, (bf16_t*)dqkv, g);
    } else {
        constexpr int TH = 8;
        dim3 grid;
        const Geom g = make_geom<TH>(B, H, W, heads, dilation, scale, &grid);
        CNB_NATC_K(ksize, CNB_NATC_HD(hd, {
            using Cfg = DkvCfg<KV, TH>;
            const size_t smem = Cfg::smem(HV);
            NATC_SET_SMEM((na_tc_bwd_dkv_kernel<HV, KV, TH>), smem);
            CNB_LAUNCH((na_tc_bwd_dkv_kernel<HV, KV, TH>), grid, dim3(Cfg::NT), smem, (cudaStream_t)stream, (const bf16_t*)qkv,
                       (const bf16_t*)dout, lse, (const float*)dvec, (bf16_t*)dqkv, g);
        }));
    }
    CNB_CHECK_LAUNCH("na_tc_bwd_kernels");
    return CNB_OK;
}


}  // namespace natc
