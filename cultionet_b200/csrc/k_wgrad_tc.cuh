// Weight gradient of the implicit-GEMM convolution on tcgen05 tensor cores.
//
//   dWp[tap][n][koff_s + c] += sum_pixels dY[p][n] * X_s[p + tap offset][c]
//
// GEMM view: D[M = 128 output channels][N = CW <= 256 source channels] over K = pixels.  Both operands are pixel-major in
// HBM, i.e. MN-major for UMMA (the M/N index is the contiguous one): a TMA box {64 ch, TWp, THp, 1} (TWp*THp = 64 pixels)
// lands as 64 rows (one per pixel = one K index) of 128 bytes with the 128B swizzle = the canonical MN-major SWIZZLE_128B
// atom stack (8 K-rows x 128 B, atoms 1024 B apart along K = SBO); wider M/N extents are further boxes LBO = 8192 B apart.
// Each CTA owns one (tap, 128-row n tile, source channel tile) and a slice of the pixel range (split-K); partial sums leave
// TMEM through tcgen05.ld and are accumulated into the fp32 packed gradient with red.global.add.
#pragma once
#ifndef CNB_EMU
#include "k_conv_tc.cuh"

namespace cnb {
namespace tc {

constexpr int WG_PIX = 64;                    // pixels (K) per pipeline stage
constexpr int WG_BOX_BYTES = WG_PIX * 128;    // one {64 ch x 64 px} box
constexpr int WG_STAGES = 4;
constexpr int WG_MAX_CTILES = 64;
constexpr int WG_MAX_BTILES = 192;            // B tiles of one launch: up to four {64 ch x 64 px} boxes each

// The pixel loop runs over the domain of the operand that is read at UNIT coordinates (U); the other operand (G) is gathered at
// s*p + q per tap through a parity sub-grid tensor map (see k_conv_tc.cuh):
//   direct convolution      (transposed = 0): U = dY (output pixels),  G = X at  s*o + (k*dil - pad)
//   transposed convolution  (transposed = 1): U = X  (input pixels),   G = dY at s*i + (k*dil - pad)
struct WgradTcParams {
    CUtensorMap tmU;
    CUtensorMap tmG[MAX_MAPS];
    int u_is_dy;
    // A "B tile" is the N side of one accumulator: up to four boxes of 64 source channels, box j = channels [c0, c0 + 64) of the
    // source gathered for tap `tap`.  Wide sources use four boxes of one tap (256 channels); sources of 64 / 128 channels put the
    // boxes of three / two TAPS side by side, so that the dY tile staged for one tap feeds several taps' accumulators (the
    // 64-channel layers staged 24 KB per 128 MMA cycles before: 145 TFLOP/s, bound by the L2 -> shared memory feed).
    int n_btiles;
    uint8_t bt_nbox[WG_MAX_BTILES];
    uint8_t bt_tap[WG_MAX_BTILES][4];
    short bt_c0[WG_MAX_BTILES][4];    // may run past src_c: zero-filled by TMA and masked in the epilogue
    int src_c, k_off;
    int ntaps;
    short tap_dy[MAX_TAPS], tap_dx[MAX_TAPS], tap_map[MAX_TAPS], tap_w[MAX_TAPS];
    int Bn, THp, TWp, tiles_h, tiles_w;  // pixel tiling of the U domain
    int N, n_tiles, Ctot;
    int splits, pt_per_split, pixel_tiles;
    int nsub, stages;  // 128-row dY sub-tiles per CTA (1 or 2, sharing the X tile) and pipeline depth
    float* dwp;
};

// MN-major SWIZZLE_128B descriptor: LBO = stride between 64-element atoms along M/N, SBO = stride between 8-row groups along K
__device__ __forceinline__ uint64_t umma_desc_mn_sw128(uint32_t smem_addr, uint32_t lbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
__device__ __forceinline__ uint32_t umma_idesc_bf16_mn(int n) {
    return (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
}

// stage = dY boxes (2 per 128-row sub-tile) followed by up to 4 X boxes; two sub-tiles share one X tile, which cuts the bytes
// staged per MMA by a third (ncu: the single-sub-tile kernel was feed-bound at 46 % tensor-pipe activity)
constexpr int WG_RING_BYTES = 192 * 1024;
constexpr int WG_SMEM_BYTES = WG_RING_BYTES + 1024 + 256;

__global__ void __launch_bounds__(NUM_THREADS, 1) wgrad_tc_kernel(const __grid_constant__ WgradTcParams p) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;
    uint8_t* base_ptr = smem_raw + (base - raw);
    const uint32_t bars = base + WG_RING_BYTES;  // full[S] empty[S] done slot
    auto full_bar = [&](int s) { return bars + 8u * s; };
    auto empty_bar = [&](int s) { return bars + 8u * (WG_STAGES + s); };
    const uint32_t done_bar = bars + 8u * (2 * WG_STAGES);
    const uint32_t slot = bars + 8u * (2 * WG_STAGES + 1);
    volatile uint32_t* slot_ptr = reinterpret_cast<volatile uint32_t*>(base_ptr + WG_RING_BYTES + 8 * (2 * WG_STAGES + 1));
    const int nsub = p.nsub, nstages = p.stages;
    const uint32_t stage_bytes = (uint32_t)(2 * nsub + 4) * WG_BOX_BYTES;
    const uint32_t tmem_cols = nsub == 2 ? 512u : 256u;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    // work item: blockIdx.x = split * items + ((tap * n_tiles + nt) * n_ctiles + ct).  The pixel split is the SLOW index: the CTAs that
    // are resident together then walk the same pixel range for every (tap, n tile, channel tile), so X and dY come from DRAM once
    // and from L2 for the other items (ncu: 4.6x DRAM re-reads with the split as the fast index).
    const int items = p.n_tiles * p.n_btiles;
    const int w = blockIdx.x % items;
    const int split = blockIdx.x / items;
    const int bt = w % p.n_btiles;
    const int nt = w / p.n_btiles;
    const int nboxes_x = p.bt_nbox[bt];
    const int cw = 64 * nboxes_x;
    const int n0 = nt * BM * nsub;
    const int pt_begin = split * p.pt_per_split;
    int pt_end = pt_begin + p.pt_per_split;
    if (pt_end > p.pixel_tiles) pt_end = p.pixel_tiles;
    const uint32_t stage_tx = (uint32_t)(2 * nsub + nboxes_x) * WG_BOX_BYTES;
    // the transposed case gathers dY per tap: its B tiles hold boxes of ONE tap (tap0)
    const int tap0 = p.bt_tap[bt][0];
    const CUtensorMap* mapDY = p.u_is_dy ? &p.tmU : &p.tmG[p.tap_map[tap0]];

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(mapDY);
        tma_prefetch_desc(p.u_is_dy ? &p.tmG[p.tap_map[tap0]] : &p.tmU);
        for (int s = 0; s < WG_STAGES; ++s) {
            mbar_init(full_bar(s), 1);
            mbar_init(empty_bar(s), 1);
        }
        mbar_init(done_bar, 1);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(slot, tmem_cols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *slot_ptr;
    CNB_PDL_SYNC();  // see conv_tc_kernel

    if (warp == 0) {
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            const int dy_oy = p.u_is_dy ? 0 : p.tap_dy[tap0], dy_ox = p.u_is_dy ? 0 : p.tap_dx[tap0];
            const CUtensorMap* box_map[4];
            int box_c0[4], box_oy[4], box_ox[4];
            for (int j = 0; j < nboxes_x; ++j) {
                const int tj = p.bt_tap[bt][j];
                box_map[j] = p.u_is_dy ? &p.tmG[p.tap_map[tj]] : &p.tmU;
                box_c0[j] = p.bt_c0[bt][j];
                box_oy[j] = p.u_is_dy ? p.tap_dy[tj] : 0;
                box_ox[j] = p.u_is_dy ? p.tap_dx[tj] : 0;
            }
            for (int pt = pt_begin; pt < pt_end; ++pt) {
                int t = pt;
                const int tw = t % p.tiles_w;
                t /= p.tiles_w;
                const int th = t % p.tiles_h;
                const int b = t / p.tiles_h;
                const int y0 = th * p.THp, x0 = tw * p.TWp;
                mbar_wait(empty_bar(stage), phase ^ 1u);
                mbar_arrive_expect_tx(full_bar(stage), stage_tx);
                const uint32_t dst = base + stage * stage_bytes;
                for (int j = 0; j < 2 * nsub; ++j)
                    tma_load_4d(dst + j * WG_BOX_BYTES, mapDY, full_bar(stage), n0 + 64 * j, x0 + dy_ox, y0 + dy_oy, b);
                for (int j = 0; j < nboxes_x; ++j)
                    tma_load_4d(dst + (2 * nsub + j) * WG_BOX_BYTES, box_map[j], full_bar(stage), box_c0[j], x0 + box_ox[j], y0 + box_oy[j], b);
                if (++stage == nstages) {
                    stage = 0;
                    phase ^= 1u;
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        if (lane == 0) {
            const uint32_t idesc = umma_idesc_bf16_mn(cw);
            int stage = 0;
            uint32_t phase = 0;
            for (int pt = pt_begin; pt < pt_end; ++pt) {
                mbar_wait(full_bar(stage), phase);
                tc_fence_after();
                const uint32_t a_addr = base + stage * stage_bytes;
                const uint32_t b_addr = a_addr + 2 * nsub * WG_BOX_BYTES;
                for (int sub = 0; sub < nsub; ++sub) {
#pragma unroll
                    for (int k = 0; k < WG_PIX / 16; ++k) {
                        // 16 pixels (K) per instruction = two 8-row groups = 2048 bytes inside every box
                        const uint64_t adesc = umma_desc_mn_sw128(a_addr + sub * 2 * WG_BOX_BYTES + k * 2048, WG_BOX_BYTES);
                        const uint64_t bdesc = umma_desc_mn_sw128(b_addr + k * 2048, WG_BOX_BYTES);
                        umma_bf16(tmem_base + (uint32_t)(sub * 256), adesc, bdesc, idesc, (pt > pt_begin || k > 0) ? 1u : 0u);
                    }
                }
                umma_commit(empty_bar(stage));
                if (++stage == nstages) {
                    stage = 0;
                    phase ^= 1u;
                }
            }
            if (pt_end > pt_begin)
                umma_commit(done_bar);
            else
                mbar_arrive(done_bar);
        }
        __syncwarp();
    } else {
        const int quad = warp & 3;          // TMEM lane quadrant (warp id % 4); warps 2..9: two per quadrant,
        const int half = (warp - 2) >> 2;   // alternating 32-column chunks
        mbar_wait(done_bar, 0);
        tc_fence_after();
        for (int sub = 0; sub < nsub; ++sub) {
            const int n = n0 + sub * BM + quad * 32 + lane;
            const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(sub * 256);
            if (pt_end > pt_begin && n0 + sub * BM < p.N) {
                for (int c = half; c < cw / 32; c += 2) {   // 32-column chunk c = half (c & 1) of box c >> 1
                    const int box = c >> 1, cin = (c & 1) * 32;
                    const int bc0 = p.bt_c0[bt][box];
                    const int cvalid = p.src_c - bc0 - cin;  // columns of this chunk that exist in the source
                    if (cvalid <= 0) continue;               // warp-uniform
                    uint32_t v[32];
                    tmem_ld32(taddr + (uint32_t)(c * 32), v);
                    if (n < p.N) {
                        float* dchunk = p.dwp + ((long)p.tap_w[p.bt_tap[bt][box]] * p.N + n) * p.Ctot + p.k_off + bc0 + cin;
                        // 16-byte aligned when Ctot, k_off and the channel offsets are multiples of 4
                        if ((reinterpret_cast<uintptr_t>(dchunk) & 15u) == 0 && cvalid >= 32) {
#pragma unroll
                            for (int j = 0; j < 32; j += 4)
                                asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dchunk + j),
                                             "f"(__uint_as_float(v[j])), "f"(__uint_as_float(v[j + 1])), "f"(__uint_as_float(v[j + 2])),
                                             "f"(__uint_as_float(v[j + 3]))
                                             : "memory");
                        } else {
#pragma unroll
                            for (int j = 0; j < 32; ++j)
                                if (j < cvalid) atomicAdd(dchunk + j, __uint_as_float(v[j]));
                        }
                    }
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, tmem_cols);
    }
}

inline bool wgrad_eligible(const cnb_wgrad_desc* d, int dtype) {
    if (dtype != CNB_BF16) return false;
    if (d->KH * d->KW > MAX_TAPS) return false;
    if (d->stride > 1 && (d->KH * d->KW > MAX_MAPS || d->stride > 4)) return false;
    if (d->Hin > 32000 || d->Win > 32000 || d->Hout > 32000 || d->Wout > 32000) return false;
    if (d->src_stride % 8 != 0 || reinterpret_cast<uintptr_t>(d->src) % 16 != 0) return false;
    if (d->dy_stride % 8 != 0 || reinterpret_cast<uintptr_t>(d->dy) % 16 != 0) return false;
    if (cnb_div_up(d->src_c, 256) * d->KH * d->KW > WG_MAX_BTILES) return false;
    return encode_tiled_fn() != nullptr;
}

inline void pick_pixel_tile(int Hv, int Wv, int* TH, int* TW) {
    long best = -1;
    for (int tw = 64; tw >= 8; tw >>= 1) {
        const int th = WG_PIX / tw;
        const long cover = (long)cnb_div_up(Wv, tw) * tw * cnb_div_up(Hv, th) * th;
        if (best < 0 || cover < best) {
            best = cover;
            *TW = tw;
            *TH = th;
        }
    }
}

// one source slice per call (mirrors cnb_conv2d_wgrad)
inline int wgrad_tc_launch(const cnb_wgrad_desc* d, cudaStream_t stream) {
    WgradTcParams p;
    memset(&p, 0, sizeof(p));
    const int s = d->stride;
    p.u_is_dy = d->transposed ? 0 : 1;
    // U = operand at unit coordinates, G = gathered operand
    const void* u_ptr = p.u_is_dy ? d->dy : d->src;
    const int u_c = p.u_is_dy ? d->N : d->src_c, u_pitch = p.u_is_dy ? d->dy_stride : d->src_stride;
    const int u_h = p.u_is_dy ? d->Hout : d->Hin, u_w = p.u_is_dy ? d->Wout : d->Win;
    const void* g_ptr = p.u_is_dy ? d->src : d->dy;
    const int g_c = p.u_is_dy ? d->src_c : d->N, g_pitch = p.u_is_dy ? d->src_stride : d->dy_stride;
    const int g_h = p.u_is_dy ? d->Hin : d->Hout, g_w = p.u_is_dy ? d->Win : d->Wout;

    pick_pixel_tile(u_h, u_w, &p.THp, &p.TWp);
    if (make_act_map(&p.tmU, u_ptr, u_c, u_w, u_h, d->B, u_pitch, p.TWp, p.THp)) return 2;
    int nt = 0;
    for (int ky = 0; ky < d->KH; ++ky)
        for (int kx = 0; kx < d->KW; ++kx) {
            const int qy = ky * d->dil - d->pad, qx = kx * d->dil - d->pad;
            if (s == 1) {
                p.tap_dy[nt] = (short)qy;
                p.tap_dx[nt] = (short)qx;
                p.tap_map[nt] = 0;
            } else {
                const int ry = pos_mod(qy, s), rx = pos_mod(qx, s);
                if (ry >= g_h || rx >= g_w) continue;  // empty parity sub-grid: this tap's gradient stays zero
                p.tap_dy[nt] = (short)floor_div(qy, s);
                p.tap_dx[nt] = (short)floor_div(qx, s);
                p.tap_map[nt] = (short)nt;
                if (make_subgrid_map(&p.tmG[nt], g_ptr, g_c, g_w, g_h, d->B, g_pitch, s, ry, rx, p.TWp, p.THp)) return 2;
            }
            p.tap_w[nt] = (short)(ky * d->KW + kx);
            ++nt;
        }
    if (s == 1 && make_act_map(&p.tmG[0], g_ptr, g_c, g_w, g_h, d->B, g_pitch, p.TWp, p.THp)) return 2;
    p.ntaps = nt;
    if (nt == 0) return 0;

    // B tiles (see WgradTcParams): narrow sources of a direct convolution share one accumulator between several taps
    int nbt = 0;
    const int boxes_per_tap = cnb_div_up(d->src_c, 64);
    static const bool group_taps = [] {
        const char* e = getenv("CNB_WGRAD_TAP_GROUPS");
        return !(e && e[0] == '0');
    }();
    if (p.u_is_dy && boxes_per_tap <= 2 && nt > 1 && group_taps) {
        const int fit = 4 / boxes_per_tap;                  // taps that fit in one 256-column accumulator
        const int ntile = cnb_div_up(nt, fit);
        const int per = cnb_div_up(nt, ntile);              // ... spread evenly: 9 taps of 64 channels -> 3 + 3 + 3
        for (int t0 = 0; t0 < nt; t0 += per) {
            int nb = 0;
            for (int t = t0; t < t0 + per && t < nt; ++t)
                for (int j = 0; j < boxes_per_tap; ++j) {
                    p.bt_tap[nbt][nb] = (uint8_t)t;
                    p.bt_c0[nbt][nb] = (short)(64 * j);
                    ++nb;
                }
            p.bt_nbox[nbt++] = (uint8_t)nb;
        }
    } else {
        for (int t = 0; t < nt; ++t)
            for (int c0 = 0; c0 < d->src_c; c0 += 256) {
                if (nbt >= WG_MAX_BTILES) return 3;
                const int rem = d->src_c - c0;
                const int nb = rem >= 256 ? 4 : cnb_div_up(rem, 64);
                for (int j = 0; j < nb; ++j) {
                    p.bt_tap[nbt][j] = (uint8_t)t;
                    p.bt_c0[nbt][j] = (short)(c0 + 64 * j);
                }
                p.bt_nbox[nbt++] = (uint8_t)nb;
            }
    }
    p.n_btiles = nbt;
    p.src_c = d->src_c;
    p.k_off = d->k_off;
    p.Bn = d->B;
    p.tiles_h = cnb_div_up(u_h, p.THp);
    p.tiles_w = cnb_div_up(u_w, p.TWp);
    p.pixel_tiles = d->B * p.tiles_h * p.tiles_w;
    p.N = d->N;
    p.nsub = d->N > BM ? 2 : 1;
    p.stages = WG_RING_BYTES / ((2 * p.nsub + 4) * WG_BOX_BYTES);
    if (p.stages > WG_STAGES) p.stages = WG_STAGES;
    p.n_tiles = cnb_div_up(d->N, BM * p.nsub);
    p.Ctot = d->Ctot;
    p.dwp = d->dwp;
    const int base_items = p.n_tiles * p.n_btiles;
    // `waves` CTAs per SM, rounded DOWN so that items * splits fills whole waves (450 CTAs on 148 SMs left the last wave 96 % empty).
    // Every CTA ends with a 128 x cw (x nsub) fp32 red.add epilogue, so the atomic traffic grows with the split count: one wave
    // instead of three took all wgrad launches of config 2 from 17.3 to 15.2 ms (782 -> 891 TFLOP/s; 32x32 shapes 490 -> 740).
    static const int waves = [] {
        const char* e = getenv("CNB_WGRAD_WAVES");
        const int w = e ? atoi(e) : 0;
        return w >= 1 && w <= 8 ? w : 1;
    }();
    int splits = (waves * num_sms()) / base_items;
    const int max_splits = p.pixel_tiles / 8 > 0 ? p.pixel_tiles / 8 : 1;
    if (splits > max_splits) splits = max_splits;
    if (splits < 1) splits = 1;
    p.pt_per_split = cnb_div_up(p.pixel_tiles, splits);
    p.splits = cnb_div_up(p.pixel_tiles, p.pt_per_split);
    static bool configured = false;
    if (!configured) {
        if (cudaFuncSetAttribute(wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, WG_SMEM_BYTES) != cudaSuccess) return 1;
        configured = true;
    }
    if (tc_pdl_enabled())
        CNB_LAUNCH(wgrad_tc_kernel, dim3(base_items * p.splits), dim3(NUM_THREADS), (size_t)WG_SMEM_BYTES, stream, p);
    else {
        cnb_count_launch();
        wgrad_tc_kernel<<<base_items * p.splits, NUM_THREADS, WG_SMEM_BYTES, stream>>>(p);
    }
    return 0;
}

}  // namespace tc
}  // namespace cnb
#endif
