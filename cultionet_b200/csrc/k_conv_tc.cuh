// Implicit-GEMM convolution on the 5th-generation tensor cores (sm_100a): tcgen05.mma with TMEM accumulators, operands
// staged by TMA (cp.async.bulk.tensor) into 128B-swizzled shared memory, warp-specialised and persistent.
//
//   D[128 pixels x BN channels] += A[128 pixels x 64 ch] * B[BN x 64 ch]^T   per (tap, source, 64-channel chunk)
//
// * A tile = a TH x TW patch of ONE image (TH*TW = 128) of a pixel-major bf16 tensor, fetched with a 4-D tensor map
//   {C, W, H, B} and box {64, TW, TH, 1} at the tap-shifted coordinate; out-of-image coordinates are zero-filled by TMA, which
//   IS the convolution's zero padding.  Box rows are 128 bytes => the K-major SWIZZLE_128B canonical UMMA layout (SBO 1024 B).
// * B tile = BN rows of the packed weights [tap][N][Ctot] (3-D map, box {64, BN, 1}), same layout.
// * the K loop walks taps x sources x chunks, so the UNet3+ concatenation is never materialised (one tensor map per source).
// * accumulators are double-buffered in TMEM (2 x BN fp32 columns): the epilogue warps drain tile i (tcgen05.ld -> +bias / BatchNorm
//   scale-shift-activation / batch-statistics sums -> bf16 -> staged row-contiguous stores) while the MMA warp already runs tile i+1.
// * two optional layouts of the ring: weights-resident (B loaded once per CTA, A-only stages) and patch mode (one TH+2-row box per
//   column shift, the taps read through row-shifted descriptors) -- see ConvTcParams.
// Warp roles (320 threads): warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer, warps 2..9 = epilogue (TMEM lane
// quadrant = warp_id % 4, two warps per quadrant alternating 32-column chunks).  One CTA per SM, static round-robin tile schedule.
#pragma once
#ifndef CNB_EMU
#include <cuda.h>

#include "cnb_common.cuh"

namespace cnb {
namespace tc {

constexpr int BM = 128;
constexpr int BK = 64;
constexpr int A_BYTES = BM * BK * 2;
constexpr int NUM_THREADS = 320;  // warp 0 TMA, warp 1 MMA, warps 2..9 epilogue (two per TMEM lane quadrant)
constexpr int NUM_EPI_WARPS = 8;
constexpr int MAX_TAPS = 16;    // taps summed over all phases
constexpr int MAX_PHASES = 16;  // sub-pixel phases of a transposed convolution (stride^2)
constexpr int MAX_MAPS = 9;     // A-operand tensor maps: one per source (unit stride) or one per tap (strided direct gather)
constexpr int EPI_CHUNK_BYTES = 32 * 64;            // 32 rows x 32 bf16 columns staged per warp per chunk
constexpr int EPI_STAGING_BYTES = NUM_EPI_WARPS * EPI_CHUNK_BYTES;  // one staging tile per epilogue warp (row-contiguous stores)
constexpr int MAX_STAGES = 12;  // barrier slots; the weights-resident mode runs up to 12 A-only stages

// The iteration space is a list of PHASES.  A phase is a unit-stride gather over a (Hv x Wv) domain per image with its own taps;
// domain pixel (vy, vx) produces output pixel (vy*osy + ooy, vx*osx + oox).
//   * unit-stride direct / transposed convolution: one phase, k*k taps, identity output map, one tensor map per source;
//   * ConvTranspose2d with stride s (and the data gradient of a stride-s convolution): s*s sub-pixel phases; phase (py, px) owns
//     the taps with (py + pad - ky*dil) % s == 0 and reads the input at domain offset (py + pad - ky*dil) / s;
//   * stride-s direct convolution (and the data gradient of a stride-s ConvTranspose2d): one phase; tap (ky, kx) reads the
//     parity sub-grid {s*i + r} of the input through its OWN tensor map (base offset r, strides x s) at domain offset floor(q/s),
//     q = ky*dil - pad, r = q mod s.
struct ConvTcParams {
    CUtensorMap tmA[MAX_MAPS];
    CUtensorMap tmB;
    int vec_ok;   // 16-byte output pitch: 128-bit stores
    int log2_tw;
    int nsrc, per_tap_map;
    int chunks[CNB_MAX_SRC];  // 64-channel chunks per source (the last one may be partial: TMA zero-fills the tail)
    int koff[CNB_MAX_SRC];    // channel offset of the source inside Ctot
    int nphases;
    int ph_tap0[MAX_PHASES + 1];   // taps of phase p: [ph_tap0[p], ph_tap0[p+1])
    int ph_tile0[MAX_PHASES + 1];  // first pixel tile of phase p
    short ph_hv[MAX_PHASES], ph_wv[MAX_PHASES], ph_tiles_h[MAX_PHASES], ph_tiles_w[MAX_PHASES], ph_ooy[MAX_PHASES], ph_oox[MAX_PHASES];
    short tap_dy[MAX_TAPS], tap_dx[MAX_TAPS], tap_w[MAX_TAPS], tap_map[MAX_TAPS];
    int TH, TW;
    int N, m_tiles, num_tiles;
    int Hout, Wout, osy, osx;
    bf16_t* out;
    int out_stride;
    const float* bias;
    const float* ep_scale;  // fused inference epilogue (cnb_conv_desc::ep_scale / ep_shift / ep_act), or NULL
    const float* ep_shift;
    int ep_act;
    float* stats;  // optional [2*N]: per-output-channel sum and sum of squares of the STORED (bf16-rounded) outputs, for BatchNorm
    int par_smem;    // bias / ep_scale / ep_shift are copied to shared memory once per CTA (N <= STATS_MAX_N)
    int staged;      // epilogue stores go through the warp's staging tile: a store instruction writes 8 rows x 64 contiguous bytes
    // weights-resident mode (one N tile, all taps x chunks of B fit beside >= 4 A stages): B is loaded ONCE per CTA into the first
    // res_bytes of the ring region and the pipeline stages carry the A operand only
    int b_resident, res_bytes, nstages;
    // Patch mode (weights-resident 3x3, stride 1, one 64-channel chunk: the 64 -> 64 layers).  Nine tap-shifted boxes of a tile overlap
    // almost completely, and ncu showed these layers waiting on TMA (0.9 GB of L2 -> SM traffic for a 67 MB input).  Here ONE box per
    // column shift dx, TH + 2 rows high, is staged (tmA[1]); the three taps that share dx read it through UMMA descriptors whose start
    // address is shifted by whole pixel rows -- TW * 128 bytes, a multiple of the 1024-byte swizzle atom for TW >= 8, so the 128-byte
    // swizzle phase of every row is unchanged.  Three TMA boxes (60 KB at TH = 8) per tile instead of nine (144 KB) plus nine B tiles.
    int patch;               // 0 = off
    int patch_bytes;         // (TH + 2 * halo) * TW * 128
    short patch_dx[3];       // column shift of group g
    short patch_y;           // row shift of the box origin (the smallest tap_dy)
    short patch_tap[3][3];   // group g, member m: absolute tap index (its resident B tile), -1 = none
    short patch_row[3][3];   // ... and its row offset inside the patch: tap_dy - patch_y
    int prefetch;  // tiles (per CTA) the producer prefetches ahead into L2; 0 = off
    int dbg;  // diagnostics only (CNB_EPI_DEBUG bit mask, timing experiments): 1 = epilogue skips its global stores, 2 = skips the BatchNorm sums
    // split output (cnb_conv_desc::nout): columns [seg_begin[i], seg_begin[i+1]) of the GEMM go to seg_out[i]; boundaries are
    // multiples of 32, so every 32-column epilogue chunk has one destination
    int nseg;
    int seg_begin[CNB_MAX_SRC + 1];
    bf16_t* seg_out[CNB_MAX_SRC];
    int seg_stride[CNB_MAX_SRC];
};

struct TileCoord {
    int ph, b, y0, x0, n0;
};

// ---------------------------------------------------------------------------------------------------------------------
// PTX wrappers
// ---------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// bounded wait: a protocol bug must trap (error returned to the host) instead of hanging the GPU
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done = 0;
    const long long t0 = clock64();
    while (true) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(bar), "r"(parity)
            : "memory");
        if (done) break;
        if (clock64() - t0 > 4000000000LL) {
            printf("cultionet_b200: mbarrier wait timed out (block %d thread %d bar %u parity %u)\n", (int)blockIdx.x, (int)threadIdx.x, bar,
                   parity);
            __trap();
        }
    }
}
__device__ __forceinline__ void tma_prefetch_desc(const void* tmap) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(tmap)) : "memory");
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const void* tmap, uint32_t bar, int c0, int c1, int c2, int c3) {
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(dst),
                 "l"(reinterpret_cast<uint64_t>(tmap)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
                 : "memory");
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const void* tmap, uint32_t bar, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(dst),
                 "l"(reinterpret_cast<uint64_t>(tmap)), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
                 : "memory");
}
// L2 prefetch of one box (no shared memory, no barrier): the producer runs it a few tiles ahead, so that the tile loads themselves are
// L2 hits.  The 1x1 / short-K convolutions keep only 64 KB of the A operand in flight per SM (four stages); at DRAM latency that is
// ~3 TB/s for the whole GPU, and any extra traffic (the epilogue's stores) stretched the latency and with it the step.
__device__ __forceinline__ void tma_prefetch_4d(const void* tmap, int c0, int c1, int c2, int c3) {
    asm volatile("cp.async.bulk.prefetch.tensor.4d.L2.global.tile [%0, {%1, %2, %3, %4}];" ::"l"(reinterpret_cast<uint64_t>(tmap)), "r"(c0),
                 "r"(c1), "r"(c2), "r"(c3)
                 : "memory");
}
__device__ __forceinline__ void tma_store_4d(const void* tmap, uint32_t src, int c0, int c1, int c2, int c3) {
    asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"(reinterpret_cast<uint64_t>(tmap)),
                 "r"(src), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
                 : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// all but the most recent bulk group have finished READING shared memory (the buffer used two chunks ago is free again)
__device__ __forceinline__ void tma_store_wait_read1() { asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tmem_alloc(uint32_t slot_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(slot_smem), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// issue only: the registers are written asynchronously and must not be read before tmem_ld_wait(v)
__device__ __forceinline__ void tmem_ld32_issue(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
          "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
          "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
          "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr)
        : "memory");
}
// wait for every outstanding tcgen05.ld of this thread.  The empty volatile statements that follow make every register an in/out
// operand of something ordered AFTER the wait (volatile asm statements keep their order), so no use of them can be scheduled above it.
__device__ __forceinline__ void tmem_ld_wait(uint32_t (&v)[32]) {
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) asm volatile("" : "+r"(v[i]));
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
          "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
          "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
          "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// K-major, SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor bit layout): 8-row groups 1024 B apart
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);  // start address            bits [0,14)
    d |= (uint64_t)1 << 16;                        // leading byte offset (unused for swizzled K-major) bits [16,30)
    d |= (uint64_t)(1024 >> 4) << 32;              // stride byte offset       bits [32,46)
    d |= (uint64_t)1 << 46;                        // descriptor version (Blackwell) bits [46,48)
    d |= (uint64_t)2 << 61;                        // SWIZZLE_128B             bits [61,64)
    return d;
}
// kind::f16 instruction descriptor: D=f32, A=B=bf16, both K-major, M=128, N=BN
__host__ __device__ constexpr uint32_t umma_idesc_bf16(int n) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
}

template <int BN>
struct Cfg {
    static constexpr int B_BYTES = BN * BK * 2;
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    static constexpr int STAGES = BN == 256 ? 4 : (BN == 128 ? 6 : 8);
    static constexpr int TMEM_COLS = 2 * BN < 32 ? 32 : 2 * BN;  // power of two for BN in {32,64,128,256}
    static constexpr int RING_BYTES = STAGES * STAGE_BYTES;
    static constexpr int SMEM_BYTES = RING_BYTES + EPI_STAGING_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;
};
constexpr int STATS_MAX_N = 1024;  // channels whose BatchNorm partial sums fit behind the barriers (8 KB)

// Column sums across the 32 lanes of a warp for 32 columns at once: lane L ends up with sum over lanes of vals[L].  Recursive
// halving: at each step a lane keeps one half of its columns and trades the other half with its partner (31 shuffles instead of
// 32 x 5 for one butterfly per column).
__device__ __forceinline__ float warp_transpose_sum32(float (&vals)[32]) {
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int step = 16; step >= 1; step >>= 1) {
        const bool upper = (lane & step) != 0;
#pragma unroll
        for (int j = 0; j < step; ++j) {
            const float send = upper ? vals[j] : vals[j + step];
            const float keep = upper ? vals[j + step] : vals[j];
            vals[j] = keep + __shfl_xor_sync(0xffffffffu, send, step);
        }
    }
    return vals[0];
}

__device__ __forceinline__ void decode_tile(const ConvTcParams& p, int tile, int bn, TileCoord& t) {
    // the output-column tile is the FAST index: the CTAs that run side by side share one pixel tile, so its A operand comes from DRAM
    // once and from L2 for the other column tiles (ncu, column tile slow: the 268 MB dY of the 256 -> 960 data gradient was read
    // 4 x from DRAM, 1.08 GB)
    const int n_tiles = p.num_tiles / p.m_tiles;
    const int nt = tile % n_tiles;
    int mt = tile / n_tiles;
    int ph = 0;
    while (ph + 1 < p.nphases && mt >= p.ph_tile0[ph + 1]) ++ph;
    mt -= p.ph_tile0[ph];
    const int tiles_w = p.ph_tiles_w[ph], tiles_h = p.ph_tiles_h[ph];
    const int tw = mt % tiles_w;
    mt /= tiles_w;
    const int th = mt % tiles_h;
    t.ph = ph;
    t.b = mt / tiles_h;
    t.y0 = th * p.TH;
    t.x0 = tw * p.TW;
    t.n0 = nt * bn;
}

template <int BN>
__global__ void __launch_bounds__(NUM_THREADS, 1) conv_tc_kernel(const __grid_constant__ ConvTcParams p) {
    using C = Cfg<BN>;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;          // SWIZZLE_128B atoms need 1024 B alignment
    uint8_t* base_ptr = smem_raw + (base - raw);
    // layout: [resident B (res_bytes) | pipeline stages]  = C::RING_BYTES, [epilogue staging], [barriers 256 B], [2*N floats]
    const uint32_t staging = base + C::RING_BYTES;
    const uint32_t bars = staging + EPI_STAGING_BYTES;  // full[MAX_STAGES] empty[MAX_STAGES] tmem_full[2] tmem_empty[2] bres slot
    auto full_bar = [&](int s) { return bars + 8u * s; };
    auto empty_bar = [&](int s) { return bars + 8u * (MAX_STAGES + s); };
    auto tfull_bar = [&](int a) { return bars + 8u * (2 * MAX_STAGES + a); };
    auto tempty_bar = [&](int a) { return bars + 8u * (2 * MAX_STAGES + 2 + a); };
    const uint32_t bres_bar = bars + 8u * (2 * MAX_STAGES + 4);
    const uint32_t slot = bars + 8u * (2 * MAX_STAGES + 5);
    volatile uint32_t* slot_ptr =
        reinterpret_cast<volatile uint32_t*>(base_ptr + C::RING_BYTES + EPI_STAGING_BYTES + 8 * (2 * MAX_STAGES + 5));
    const int nstages = p.nstages;
    const uint32_t stage_bytes = p.patch ? (uint32_t)p.patch_bytes : (p.b_resident ? (uint32_t)A_BYTES : (uint32_t)C::STAGE_BYTES);
    const uint32_t ring0 = base + (uint32_t)p.res_bytes;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // [2*N] floats: BatchNorm partial sums (p.stats), or bias / scale | shift of the fused epilogue (p.par_smem)
    float* sm_par = reinterpret_cast<float*>(base_ptr + C::RING_BYTES + EPI_STAGING_BYTES + 256);
    if (p.stats)
        for (int i = threadIdx.x; i < 2 * p.N; i += NUM_THREADS) sm_par[i] = 0.f;

    if (warp == 0 && lane == 0) {
        const int nmaps = p.per_tap_map ? p.ph_tap0[p.nphases] : p.nsrc;
        for (int s = 0; s < nmaps; ++s) tma_prefetch_desc(&p.tmA[s]);
        tma_prefetch_desc(&p.tmB);
        for (int s = 0; s < nstages; ++s) {
            mbar_init(full_bar(s), 1);
            mbar_init(empty_bar(s), 1);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(tfull_bar(a), 1);
            mbar_init(tempty_bar(a), NUM_EPI_WARPS);  // one arrive per epilogue warp
        }
        mbar_init(bres_bar, 1);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(slot, C::TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *slot_ptr;
    // programmatic dependent launch: everything above (barriers, TMEM, descriptor prefetch) ran while the previous kernel of the
    // stream was draining; nothing below may touch global memory before that kernel has completed
    CNB_PDL_SYNC();

    int chunks_per_tap = 0;
    for (int s = 0; s < p.nsrc; ++s) chunks_per_tap += p.chunks[s];

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            if (p.b_resident) {
                // every (tap, source, chunk) tile of the packed weights, once per CTA, in the order the K loop walks them
                mbar_arrive_expect_tx(bres_bar, (uint32_t)p.res_bytes);
                uint32_t dst = base;
                for (int t = 0; t < p.ph_tap0[p.nphases]; ++t)
                    for (int s = 0; s < p.nsrc; ++s)
                        for (int kc = 0; kc < p.chunks[s]; ++kc) {
                            tma_load_3d(dst, &p.tmB, bres_bar, p.koff[s] + kc * BK, 0, p.tap_w[t]);
                            dst += C::B_BYTES;
                        }
            }
            int stage = 0;
            uint32_t phase = 0;
            TileCoord tc;
            // L2 prefetch of the un-shifted box of a pixel tile, per source chunk (the halo comes from the neighbouring tiles' boxes,
            // which other CTAs prefetch at about the same time); only the first N tile of a pixel tile does it
            auto prefetch_tile = [&](int tile) {
                if (tile >= p.num_tiles || p.per_tap_map) return;
                TileCoord pc;
                decode_tile(p, tile, BN, pc);
                if (pc.n0 != 0 || p.ph_tap0[pc.ph + 1] == p.ph_tap0[pc.ph]) return;
                int t0 = p.ph_tap0[pc.ph];  // the tap with the smallest shift
                for (int t = t0 + 1; t < p.ph_tap0[pc.ph + 1]; ++t)
                    if (abs(p.tap_dy[t]) + abs(p.tap_dx[t]) < abs(p.tap_dy[t0]) + abs(p.tap_dx[t0])) t0 = t;
                for (int s = 0; s < p.nsrc; ++s)
                    for (int kc = 0; kc < p.chunks[s]; ++kc)
                        tma_prefetch_4d(&p.tmA[s], kc * BK, pc.x0 + p.tap_dx[t0], pc.y0 + p.tap_dy[t0], pc.b);
            };
            if (p.prefetch)
                for (int a = 1; a <= p.prefetch; ++a) prefetch_tile(blockIdx.x + a * (int)gridDim.x);
            for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
                decode_tile(p, tile, BN, tc);
                if (p.prefetch) prefetch_tile(tile + (p.prefetch + 1) * (int)gridDim.x);
                if (p.patch) {
                    for (int gq = 0; gq < 3; ++gq) {
                        mbar_wait(empty_bar(stage), phase ^ 1u);
                        mbar_arrive_expect_tx(full_bar(stage), (uint32_t)p.patch_bytes);
                        tma_load_4d(ring0 + stage * stage_bytes, &p.tmA[1], full_bar(stage), 0, tc.x0 + p.patch_dx[gq], tc.y0 + p.patch_y, tc.b);
                        if (++stage == nstages) {
                            stage = 0;
                            phase ^= 1u;
                        }
                    }
                    continue;
                }
                for (int t = p.ph_tap0[tc.ph]; t < p.ph_tap0[tc.ph + 1]; ++t) {
                    const int cy = tc.y0 + p.tap_dy[t], cx = tc.x0 + p.tap_dx[t], wt = p.tap_w[t];
                    for (int s = 0; s < p.nsrc; ++s) {
                        const CUtensorMap* am = &p.tmA[p.per_tap_map ? p.tap_map[t] : s];
                        for (int kc = 0; kc < p.chunks[s]; ++kc) {
                            mbar_wait(empty_bar(stage), phase ^ 1u);
                            mbar_arrive_expect_tx(full_bar(stage), stage_bytes);
                            const uint32_t a_dst = ring0 + stage * stage_bytes;
                            tma_load_4d(a_dst, am, full_bar(stage), kc * BK, cx, cy, tc.b);
                            if (!p.b_resident) tma_load_3d(a_dst + A_BYTES, &p.tmB, full_bar(stage), p.koff[s] + kc * BK, tc.n0, wt);
                            if (++stage == nstages) {
                                stage = 0;
                                phase ^= 1u;
                            }
                        }
                    }
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        if (lane == 0) {
            constexpr uint32_t idesc = umma_idesc_bf16(BN);
            int stage = 0;
            uint32_t phase = 0;
            int it = 0;
            TileCoord tc;
            if (p.b_resident) {
                mbar_wait(bres_bar, 0u);
                tc_fence_after();
            }
            for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++it) {
                decode_tile(p, tile, BN, tc);
                const int total_chunks = (p.ph_tap0[tc.ph + 1] - p.ph_tap0[tc.ph]) * chunks_per_tap;
                const uint32_t b_res0 = base + (uint32_t)(p.ph_tap0[tc.ph] * chunks_per_tap) * (uint32_t)C::B_BYTES;
                const int acc = it & 1;
                const uint32_t acc_phase = (uint32_t)(it >> 1) & 1u;
                mbar_wait(tempty_bar(acc), acc_phase ^ 1u);  // epilogue has drained this accumulator
                tc_fence_after();
                const uint32_t tmem_d = tmem_base + (uint32_t)(acc * BN);
                if (p.patch) {
                    uint32_t accum = 0u;  // the first MMA of the tile overwrites the accumulator
                    for (int gq = 0; gq < 3; ++gq) {
                        mbar_wait(full_bar(stage), phase);
                        tc_fence_after();
                        const uint32_t patch_addr = ring0 + stage * stage_bytes;
                        for (int m = 0; m < 3; ++m) {
                            const int t = p.patch_tap[gq][m];
                            if (t < 0) continue;
                            const uint64_t adesc = umma_desc_sw128(patch_addr + (uint32_t)(p.patch_row[gq][m] * p.TW) * 128u);
                            const uint64_t bdesc = umma_desc_sw128(base + (uint32_t)t * (uint32_t)C::B_BYTES);
#pragma unroll
                            for (int k = 0; k < BK / 16; ++k) {
                                umma_bf16(tmem_d, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, accum);
                                accum = 1u;
                            }
                        }
                        umma_commit(empty_bar(stage));
                        if (++stage == nstages) {
                            stage = 0;
                            phase ^= 1u;
                        }
                    }
                    umma_commit(tfull_bar(acc));
                    continue;
                }
                for (int c = 0; c < total_chunks; ++c) {
                    mbar_wait(full_bar(stage), phase);
                    tc_fence_after();
                    const uint32_t a_addr = ring0 + stage * stage_bytes;
                    const uint64_t adesc = umma_desc_sw128(a_addr);
                    const uint64_t bdesc = umma_desc_sw128(p.b_resident ? b_res0 + (uint32_t)c * (uint32_t)C::B_BYTES : a_addr + A_BYTES);
#pragma unroll
                    for (int k = 0; k < BK / 16; ++k) {
                        // +32 bytes per UMMA_K step inside the 128-byte swizzled row => +2 in the (addr >> 4) field
                        umma_bf16(tmem_d, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, (c > 0 || k > 0) ? 1u : 0u);
                    }
                    umma_commit(empty_bar(stage));  // frees the smem slot when these MMAs have read it
                    if (++stage == nstages) {
                        stage = 0;
                        phase ^= 1u;
                    }
                }
                if (total_chunks > 0)
                    umma_commit(tfull_bar(acc));  // accumulator complete -> epilogue
                else
                    mbar_arrive(tfull_bar(acc));  // tap-less phase (bias only): nothing in flight, hand over directly
            }
        }
        __syncwarp();
    } else {
        // ===================== epilogue: TMEM -> registers -> bf16 -> global =====================
        // A TMEM lane holds one pixel row: after tcgen05.ld a lane owns 32 columns (64 bytes) of ITS row.  Round 2 (ncu of the 1x1 and
        // short-K convolutions, whose K loop cannot hide the epilogue: 5.5-6 us per 128 x 256 tile against 1.2 us of MMA):
        //   * bias / scale / shift come from shared memory (copied once per CTA) -- the 32-64 uniform __ldg per chunk were the top
        //     long-scoreboard and LSU-queue stalls;
        //   * the next chunk's tcgen05.ld is in flight while the current one is converted and stored (two register sets);
        //   * stores go through a 2 KB staging tile per warp so that an instruction writes 8 rows x 64 contiguous bytes (8 lines)
        //     instead of 32 rows x 16 bytes (32 lines: 32 L1 tag cycles per instruction = 2.4 us per tile by itself).
        // Eight warps, two per TMEM lane quadrant, alternate 32-column chunks.
        const int ew = warp - 2;
        const int quad = warp & 3;   // TMEM lane quadrant this warp may access (warp id % 4)
        const int half = ew >> 2;    // which of the two warps of that quadrant
        const int row = quad * 32 + lane;
        const int ly = row >> p.log2_tw, lx = row & (p.TW - 1);
        uint8_t* stg = base_ptr + C::RING_BYTES + ew * EPI_CHUNK_BYTES;
        const int sw_w = (lane >> 1) & 3;  // 16-byte chunk swizzle of this lane's own row (conflict-free 128-bit accesses both ways)
        const int sw_r = (lane >> 3) & 3;  // ... of the rows (lane >> 2) + 8 k this lane stores (the same for every k)
        const int rq = lane & 3;
        if (p.par_smem) {
            // epilogue threads only (named barrier 1): bias -> [0, N), or scale -> [0, N) and shift -> [N, 2N)
            const float* a0 = p.ep_scale ? p.ep_scale : p.bias;
            for (int i = (int)threadIdx.x - 64; i < p.N; i += NUM_EPI_WARPS * 32) {
                sm_par[i] = a0[i];
                if (p.ep_scale) sm_par[p.N + i] = p.ep_shift[i];
            }
            asm volatile("bar.sync 1, %0;" ::"n"(NUM_EPI_WARPS * 32) : "memory");
        }
        int it = 0;
        TileCoord tc;
        for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++it) {
            decode_tile(p, tile, BN, tc);
            const int acc = it & 1;
            const uint32_t acc_phase = (uint32_t)(it >> 1) & 1u;
            const int vy = tc.y0 + ly, vx = tc.x0 + lx;
            const int oy = vy * p.osy + p.ph_ooy[tc.ph], ox = vx * p.osx + p.ph_oox[tc.ph];
            const bool valid = vy < p.ph_hv[tc.ph] && vx < p.ph_wv[tc.ph] && oy < p.Hout && ox < p.Wout;
            const bool has_taps = p.ph_tap0[tc.ph + 1] > p.ph_tap0[tc.ph];
            const int n0 = tc.n0;
            const long opix = ((long)tc.b * p.Hout + oy) * p.Wout + ox;
            // rows this lane stores in staged mode: (lane >> 2) + 8 k
            long opix_k[4];
            bool valid_k[4];
            if (p.staged) {
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const int src = (lane >> 2) + 8 * k;
                    opix_k[k] = __shfl_sync(0xffffffffu, (long long)opix, src);
                    valid_k[k] = __shfl_sync(0xffffffffu, (int)valid, src) != 0;
                }
            }
            int nch = (p.N - n0 + 31) / 32;  // 32-column chunks of this tile that hold output channels
            if (nch > BN / 32) nch = BN / 32;
            mbar_wait(tfull_bar(acc), acc_phase);
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(acc * BN);

            auto finish_chunk = [&](uint32_t(&v)[32], int c) {
                const int col0 = n0 + c * 32;
                const bool full = p.vec_ok && col0 + 32 <= p.N;
                uint32_t packed[16];
                if (p.stats) {
                    // BatchNorm batch statistics of this chunk (training-mode ConvBlock2d: no bias, no fused epilogue, N % 32 == 0):
                    // per-channel sum and sum of squares of the STORED (bf16-rounded) values over the warp's 32 pixel rows -- one
                    // transposing reduction per quantity (31 shuffles), then a shared-memory atomic per lane; the CTA flushes its
                    // [2, N] partials to global memory once, at the end.  The separate statistics pass (a full read of the output) goes.
                    float vals[32];  // one array, reduced twice (values, then squares): the two register sets of the pipelined
                                     // tcgen05.ld leave no room for both at once
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        packed[j] = (has_taps && valid) ? cnb_pack_bf16x2(__uint_as_float(v[2 * j]), __uint_as_float(v[2 * j + 1])) : 0u;
                        vals[2 * j] = cnb_bits2f(packed[j] << 16), vals[2 * j + 1] = cnb_bits2f(packed[j] & 0xffff0000u);
                    }
                    if (!(p.dbg & 2)) {
                        const float s1 = warp_transpose_sum32(vals);
#pragma unroll
                        for (int j = 0; j < 16; ++j) {
                            const float f0 = cnb_bits2f(packed[j] << 16), f1 = cnb_bits2f(packed[j] & 0xffff0000u);
                            vals[2 * j] = f0 * f0, vals[2 * j + 1] = f1 * f1;
                        }
                        const float s2 = warp_transpose_sum32(vals);
                        atomicAdd(&sm_par[col0 + lane], s1);
                        atomicAdd(&sm_par[p.N + col0 + lane], s2);
                    }
                } else if (!full) {
                    // ragged chunk (partial N tile or an unaligned pitch): scalar stores
                    if (!valid) return;
                    bf16_t* ochunk = p.out + opix * p.out_stride + col0;
                    if (p.nseg) {
                        int sg = 0;
                        while (sg + 1 < p.nseg && col0 >= p.seg_begin[sg + 1]) ++sg;
                        ochunk = p.seg_out[sg] + opix * p.seg_stride[sg] + (col0 - p.seg_begin[sg]);
                    }
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        if (col0 + j < p.N) {
                            float f = has_taps ? __uint_as_float(v[j]) : 0.f;
                            if (p.bias) f += __ldg(p.bias + col0 + j);
                            if (p.ep_scale) {
                                f = fmaf(f, __ldg(p.ep_scale + col0 + j), __ldg(p.ep_shift + col0 + j));
                                f = cnb_act_t<bf16_t>(f, p.ep_act);
                            }
                            ochunk[j] = __float2bfloat16(f);
                        }
                    }
                    return;
                } else if (p.ep_scale) {
                    // eval-mode BatchNorm (+ SiLU) on the fp32 accumulator: the convolution output never exists un-normalised
                    if (p.par_smem) {
                        const float4* sc4 = reinterpret_cast<const float4*>(sm_par + col0);
                        const float4* sh4 = reinterpret_cast<const float4*>(sm_par + p.N + col0);
#pragma unroll
                        for (int q = 0; q < 8; ++q) {
                            const float4 sc = sc4[q], sh = sh4[q];
                            float f0 = fmaf(has_taps ? __uint_as_float(v[4 * q]) : 0.f, sc.x, sh.x);
                            float f1 = fmaf(has_taps ? __uint_as_float(v[4 * q + 1]) : 0.f, sc.y, sh.y);
                            float f2 = fmaf(has_taps ? __uint_as_float(v[4 * q + 2]) : 0.f, sc.z, sh.z);
                            float f3 = fmaf(has_taps ? __uint_as_float(v[4 * q + 3]) : 0.f, sc.w, sh.w);
                            if (p.ep_act)
                                f0 = cnb_act_t<bf16_t>(f0, p.ep_act), f1 = cnb_act_t<bf16_t>(f1, p.ep_act), f2 = cnb_act_t<bf16_t>(f2, p.ep_act),
                                f3 = cnb_act_t<bf16_t>(f3, p.ep_act);
                            packed[2 * q] = cnb_pack_bf16x2(f0, f1);
                            packed[2 * q + 1] = cnb_pack_bf16x2(f2, f3);
                        }
                    } else {
#pragma unroll
                        for (int j = 0; j < 16; ++j) {
                            float f0 = fmaf(has_taps ? __uint_as_float(v[2 * j]) : 0.f, __ldg(p.ep_scale + col0 + 2 * j), __ldg(p.ep_shift + col0 + 2 * j));
                            float f1 = fmaf(has_taps ? __uint_as_float(v[2 * j + 1]) : 0.f, __ldg(p.ep_scale + col0 + 2 * j + 1),
                                            __ldg(p.ep_shift + col0 + 2 * j + 1));
                            if (p.ep_act) f0 = cnb_act_t<bf16_t>(f0, p.ep_act), f1 = cnb_act_t<bf16_t>(f1, p.ep_act);
                            packed[j] = cnb_pack_bf16x2(f0, f1);
                        }
                    }
                } else if (p.bias) {
                    if (p.par_smem) {
                        const float4* b4 = reinterpret_cast<const float4*>(sm_par + col0);
#pragma unroll
                        for (int q = 0; q < 8; ++q) {
                            const float4 b = b4[q];
                            packed[2 * q] = cnb_pack_bf16x2((has_taps ? __uint_as_float(v[4 * q]) : 0.f) + b.x,
                                                            (has_taps ? __uint_as_float(v[4 * q + 1]) : 0.f) + b.y);
                            packed[2 * q + 1] = cnb_pack_bf16x2((has_taps ? __uint_as_float(v[4 * q + 2]) : 0.f) + b.z,
                                                                (has_taps ? __uint_as_float(v[4 * q + 3]) : 0.f) + b.w);
                        }
                    } else {
#pragma unroll
                        for (int j = 0; j < 16; ++j) {
                            const float f0 = (has_taps ? __uint_as_float(v[2 * j]) : 0.f) + __ldg(p.bias + col0 + 2 * j);
                            const float f1 = (has_taps ? __uint_as_float(v[2 * j + 1]) : 0.f) + __ldg(p.bias + col0 + 2 * j + 1);
                            packed[j] = cnb_pack_bf16x2(f0, f1);
                        }
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < 16; ++j)
                        packed[j] = has_taps ? cnb_pack_bf16x2(__uint_as_float(v[2 * j]), __uint_as_float(v[2 * j + 1])) : 0u;
                }
                // ---- store the chunk: 64 bytes per pixel row ----
                if (p.dbg & 1) return;
                int sg = 0;
                if (p.nseg)
                    while (sg + 1 < p.nseg && col0 >= p.seg_begin[sg + 1]) ++sg;
                bf16_t* const obase = p.nseg ? p.seg_out[sg] + (col0 - p.seg_begin[sg]) : p.out + col0;
                const long opitch = p.nseg ? p.seg_stride[sg] : p.out_stride;
                if (p.staged) {
                    uint4* wrow = reinterpret_cast<uint4*>(stg + lane * 64);
#pragma unroll
                    for (int q = 0; q < 4; ++q) wrow[q ^ sw_w] = make_uint4(packed[4 * q], packed[4 * q + 1], packed[4 * q + 2], packed[4 * q + 3]);
                    __syncwarp();
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const int rr = (lane >> 2) + 8 * k;
                        const uint4 val = *reinterpret_cast<const uint4*>(stg + rr * 64 + ((rq ^ sw_r) << 4));
                        if (valid_k[k]) *reinterpret_cast<uint4*>(obase + opix_k[k] * opitch + rq * 8) = val;
                    }
                    __syncwarp();
                } else if (valid) {
                    uint4* dst = reinterpret_cast<uint4*>(obase + opix * opitch);
#pragma unroll
                    for (int q = 0; q < 4; ++q) dst[q] = make_uint4(packed[4 * q], packed[4 * q + 1], packed[4 * q + 2], packed[4 * q + 3]);
                }
            };

            // two register sets: chunk c + 2 is on its way out of TMEM while chunk c is converted and stored
            uint32_t va[32], vb[32];
            int c = half;
            if (c < nch) tmem_ld32_issue(taddr + (uint32_t)(c * 32), va);
            while (c < nch) {
                tmem_ld_wait(va);
                if (c + 2 < nch) tmem_ld32_issue(taddr + (uint32_t)((c + 2) * 32), vb);
                finish_chunk(va, c);
                c += 2;
                if (c >= nch) break;
                tmem_ld_wait(vb);
                if (c + 2 < nch) tmem_ld32_issue(taddr + (uint32_t)((c + 2) * 32), va);
                finish_chunk(vb, c);
                c += 2;
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tempty_bar(acc));
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, C::TMEM_COLS);
    }
    if (p.stats)
        for (int i = threadIdx.x; i < 2 * p.N; i += NUM_THREADS) {
            const float t = sm_par[i];
            if (t != 0.f) atomicAdd(p.stats + i, t);
        }
}

// ---------------------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline EncodeTiledFn encode_tiled_fn() {
    static EncodeTiledFn fn = [] {
        void* f = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess)
            f = nullptr;
        return reinterpret_cast<EncodeTiledFn>(f);
    }();
    return fn;
}

// what the last failed tensor-map encode was asked for (appended to the C ABI's error message)
inline char* encode_diag() {
    static thread_local char buf[256] = {0};
    return buf;
}

inline int num_sms() {
    static int n = [] {
        int dev = 0, v = CNB_NUM_SMS;
        if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev);
        return v > 0 ? v : CNB_NUM_SMS;
    }();
    return n;
}

// bf16 tensor map over a strided pixel grid: element (c, x, y, b) at ptr + c + x*sx + y*sy + b*sb (strides in elements);
// box {64, TW, TH, 1}; coordinates outside [0,C) x [0,W) x [0,H) read as zero
inline int make_grid_map(CUtensorMap* m, const void* ptr, int C, int W, int H, int B, long sx, long sy, long sb, int TW, int TH) {
    EncodeTiledFn fn = encode_tiled_fn();
    if (!fn) return 1;
    if (C <= 0 || W <= 0 || H <= 0 || B <= 0) return 3;
    cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
    cuuint64_t strides[3] = {(cuuint64_t)sx * 2, (cuuint64_t)sy * 2, (cuuint64_t)sb * 2};
    cuuint32_t box[4] = {(cuuint32_t)BK, (cuuint32_t)TW, (cuuint32_t)TH, 1};
    cuuint32_t es[4] = {1, 1, 1, 1};
    CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(ptr), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r == CUDA_ERROR_INVALID_CONTEXT || r == CUDA_ERROR_NOT_INITIALIZED) {
        // a driver-API call on a thread that has made no runtime call yet (autograd's backward thread when every allocation came out
        // of torch's cache): bind the primary context of the current device to this thread and encode again
        cudaFree(nullptr);
        r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(ptr), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    }
    if (r != CUDA_SUCCESS)
        snprintf(encode_diag(), 256, "grid map: driver code %d, ptr %p, dims {%d,%d,%d,%d}, strides {%ld,%ld,%ld} elements, box {%d,%d,%d}", (int)r,
                 ptr, C, W, H, B, sx, sy, sb, BK, TW, TH);
    return r == CUDA_SUCCESS ? 0 : 2;
}

// dense pixel-major activation [B][H][W][stride_px] taking C channels
inline int make_act_map(CUtensorMap* m, const void* ptr, int C, int W, int H, int B, int stride_px, int TW, int TH) {
    return make_grid_map(m, ptr, C, W, H, B, stride_px, (long)W * stride_px, (long)H * W * stride_px, TW, TH);
}

// parity sub-grid {(s*y + ry, s*x + rx)} of a dense activation
inline int make_subgrid_map(CUtensorMap* m, const void* ptr, int C, int W, int H, int B, int stride_px, int s, int ry, int rx, int TW,
                            int TH) {
    const int Hs = (H - ry + s - 1) / s, Ws = (W - rx + s - 1) / s;
    const bf16_t* base = reinterpret_cast<const bf16_t*>(ptr) + ((long)ry * W + rx) * stride_px;
    return make_grid_map(m, base, C, Ws, Hs, B, (long)s * stride_px, (long)s * W * stride_px, (long)H * W * stride_px, TW, TH);
}

// bf16 tensor map over packed weights [taps][rows][row_stride] taking K columns; box {64, BN, 1}
inline int make_weight_map(CUtensorMap* m, const void* ptr, int K, int rows, int taps, long row_stride, long tap_stride, int BN) {
    EncodeTiledFn fn = encode_tiled_fn();
    if (!fn) return 1;
    cuuint64_t dims[3] = {(cuuint64_t)K, (cuuint64_t)rows, (cuuint64_t)taps};
    cuuint64_t strides[2] = {(cuuint64_t)row_stride * 2, (cuuint64_t)tap_stride * 2};
    cuuint32_t box[3] = {(cuuint32_t)BK, (cuuint32_t)BN, 1};
    cuuint32_t es[3] = {1, 1, 1};
    CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(ptr), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r == CUDA_ERROR_INVALID_CONTEXT || r == CUDA_ERROR_NOT_INITIALIZED) {
        cudaFree(nullptr);  // see make_grid_map
        r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(ptr), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    }
    if (r != CUDA_SUCCESS)
        snprintf(encode_diag(), 256, "weight map: driver code %d, ptr %p, dims {%d,%d,%d}, strides {%ld,%ld} elements, box {%d,%d}", (int)r, ptr, K,
                 rows, taps, row_stride, tap_stride, BK, BN);
    return r == CUDA_SUCCESS ? 0 : 2;
}

inline void pick_tile(int Hv, int Wv, int* TH, int* TW) {
    long best = -1;
    for (int tw = 128; tw >= 8; tw >>= 1) {
        const int th = BM / tw;
        const long cover = (long)cnb_div_up(Wv, tw) * tw * cnb_div_up(Hv, th) * th;
        if (best < 0 || cover < best) {
            best = cover;
            *TW = tw;
            *TH = th;
        }
    }
}

static inline int floor_div(int a, int b) { return (a >= 0) ? a / b : -((-a + b - 1) / b); }
static inline int pos_mod(int a, int b) { return a - floor_div(a, b) * b; }

// Is this convolution one the tensor-core kernel takes?  bf16; every source pixel pitch a multiple of 8 channels (16 bytes) and
// 16-byte aligned; strided / stride-transposed gathers need a single source and at most 9 taps.  Channel counts are free: partial
// 64-channel chunks and partial N tiles are zero-filled by TMA and masked in the epilogue.
inline bool eligible(const cnb_conv_desc* d, int dtype) {
    if (dtype != CNB_BF16) return false;
    if (d->KH * d->KW > MAX_TAPS) return false;
    if (d->stride > 1 && (d->nsrc != 1 || d->KH * d->KW > MAX_MAPS || d->stride > 4)) return false;
    if (d->Hin > 32000 || d->Win > 32000 || d->Hout > 32000 || d->Wout > 32000) return false;
    for (int s = 0; s < d->nsrc; ++s) {
        if (d->src_stride[s] % 8 != 0) return false;
        if (reinterpret_cast<uintptr_t>(d->src[s]) % 16 != 0) return false;
    }
    if (d->w_row_stride % 8 != 0 || d->w_tap_stride % 8 != 0 || reinterpret_cast<uintptr_t>(d->w_packed) % 16 != 0) return false;
    for (int i = 0; i < d->nout; ++i)
        if (d->out_seg_c[i] % 32 != 0 || d->out_seg_stride[i] % 8 != 0 || reinterpret_cast<uintptr_t>(d->out_seg[i]) % 16 != 0) return false;
    return encode_tiled_fn() != nullptr;
}

// BatchNorm statistics from the epilogue (cnb_conv_desc::stats): plain single-destination convolution without bias / fused epilogue whose
// column count is a whole number of 32-column chunks (every ConvBlock2d of the TowerUNet); CNB_CONV_STATS=0 switches it off (A/B)
inline bool stats_ok(const cnb_conv_desc* d) {
    static const bool on = [] {
        const char* e = getenv("CNB_CONV_STATS");
        return !(e && e[0] == '0');
    }();
    return on && d->N % 32 == 0 && d->N <= STATS_MAX_N && !d->bias && !d->ep_scale && d->nout == 0 && d->out && d->out_stride % 8 == 0 &&
           reinterpret_cast<uintptr_t>(d->out) % 16 == 0;
}

// CNB_PDL_TC=0 launches the tensor-core kernels without the programmatic-dependent-launch attribute (A/B timing)
inline bool tc_pdl_enabled() {
    static const bool on = [] {
        const char* e = getenv("CNB_PDL_TC");
        return !(e && e[0] == '0');
    }();
    return on;
}

// A/B switches of the round-2 epilogue and of the weights-resident mode (default on): CNB_EPI_STAGED=0, CNB_EPI_PARSMEM=0, CNB_B_RESIDENT=0
inline bool env_on(const char* name) {
    const char* e = getenv(name);
    return !(e && e[0] == '0');
}

// pipeline depth / weights-resident decision (see ConvTcParams): resident when the launch has ONE N tile, every (tap, source, chunk)
// tile of B fits beside at least four A-only stages, and a CTA runs enough pixel tiles to pay for loading all of B up front
template <int BN>
inline void plan_ring(ConvTcParams& p, int total_b_tiles) {
    static const bool allow = env_on("CNB_B_RESIDENT");
    p.b_resident = 0, p.res_bytes = 0, p.nstages = Cfg<BN>::STAGES;
    const long res = (long)total_b_tiles * Cfg<BN>::B_BYTES;
    if (p.patch) {  // conv_tc_launch has checked that B and three patches fit and that the launch is big enough
        int stages = (int)((Cfg<BN>::RING_BYTES - res) / p.patch_bytes);
        p.b_resident = 1;
        p.res_bytes = (int)res;
        p.nstages = stages < MAX_STAGES ? stages : MAX_STAGES;
        return;
    }
    if (!allow || p.num_tiles != p.m_tiles || total_b_tiles <= 0 || res > Cfg<BN>::RING_BYTES) return;
    const int stages = (int)((Cfg<BN>::RING_BYTES - res) / A_BYTES);
    if (stages < 4 || p.m_tiles < 4 * num_sms()) return;
    p.b_resident = 1;
    p.res_bytes = (int)res;
    p.nstages = stages < MAX_STAGES ? stages : MAX_STAGES;
}

template <int BN>
inline int launch_bn(ConvTcParams& p, cudaStream_t stream) {
    {
        int cpt = 0;
        for (int s = 0; s < p.nsrc; ++s) cpt += p.chunks[s];
        plan_ring<BN>(p, p.ph_tap0[p.nphases] * cpt);
    }
    static bool configured = false;
    if (!configured) {
        if (cudaFuncSetAttribute(conv_tc_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 Cfg<BN>::SMEM_BYTES + 2 * STATS_MAX_N * (int)sizeof(float)) != cudaSuccess)
            return 1;
        configured = true;
    }
    const int grid = p.num_tiles < num_sms() ? p.num_tiles : num_sms();
    const int smem = Cfg<BN>::SMEM_BYTES + ((p.stats || p.par_smem) ? 2 * p.N * (int)sizeof(float) : 0);
    if (tc_pdl_enabled())
        CNB_LAUNCH(conv_tc_kernel<BN>, dim3(grid), dim3(NUM_THREADS), (size_t)smem, stream, p);
    else {
        cnb_count_launch();
        conv_tc_kernel<BN><<<grid, NUM_THREADS, smem, stream>>>(p);
    }
    return 0;
}

// Output-tile width: the padded column count times a penalty for narrow tiles (a narrow tile re-stages the A operand more often per
// output column).  N = 960 (merged data gradient of a tower convolution) takes 4 tiles of 256 with the last one masked, not 15 of 64.
inline int pick_bn(int N) {
    static const int forced = [] {  // CNB_TC_BN=32|64|128|256: one tile width for every launch (timing experiments)
        const char* e = getenv("CNB_TC_BN");
        const int v = e ? atoi(e) : 0;
        return (v == 32 || v == 64 || v == 128 || v == 256) ? v : 0;
    }();
    if (forced) return forced;
    if (N <= 32) return 32;
    const int bns[4] = {256, 128, 64, 32};
    const float penalty[4] = {1.0f, 1.15f, 1.6f, 2.5f};
    int best = 256;
    float best_cost = 1e30f;
    for (int i = 0; i < 4; ++i) {
        const float cost = (float)(cnb_div_up(N, bns[i]) * bns[i]) * penalty[i];
        if (cost < best_cost) best_cost = cost, best = bns[i];
    }
    return best;
}

inline int conv_tc_launch(const cnb_conv_desc* d, cudaStream_t stream) {
    ConvTcParams p;
    memset(&p, 0, sizeof(p));
    encode_diag()[0] = 0;
    const int BN = pick_bn(d->N);
    const int s = d->stride;
    const int taps = d->KH * d->KW;

    // ---- phases, taps, output map ----
    int dom_h, dom_w;  // largest per-image iteration domain
    if (s == 1 || !d->transposed) {
        p.nphases = 1;
        p.ph_tap0[0] = 0;
        p.ph_hv[0] = (short)d->Hout;
        p.ph_wv[0] = (short)d->Wout;
        p.osy = p.osx = 1;
        dom_h = d->Hout;
        dom_w = d->Wout;
        int nt = 0;
        for (int ky = 0; ky < d->KH; ++ky)
            for (int kx = 0; kx < d->KW; ++kx) {
                const int t = ky * d->KW + kx;
                if (s == 1) {
                    p.tap_dy[nt] = (short)(d->transposed ? d->pad - ky * d->dil : ky * d->dil - d->pad);
                    p.tap_dx[nt] = (short)(d->transposed ? d->pad - kx * d->dil : kx * d->dil - d->pad);
                    p.tap_map[nt] = 0;
                } else {
                    const int qy = ky * d->dil - d->pad, qx = kx * d->dil - d->pad;
                    const int ry = pos_mod(qy, s), rx = pos_mod(qx, s);
                    if (ry >= d->Hin || rx >= d->Win) continue;  // the parity sub-grid is empty: the tap never lands inside
                    p.tap_dy[nt] = (short)floor_div(qy, s);
                    p.tap_dx[nt] = (short)floor_div(qx, s);
                    p.tap_map[nt] = (short)nt;
                }
                p.tap_w[nt] = (short)t;
                ++nt;
            }
        p.ph_tap0[1] = nt;
        p.per_tap_map = s > 1;
    } else {
        // transposed, stride s: sub-pixel phases
        p.osy = p.osx = s;
        int np = 0, nt = 0;
        dom_h = dom_w = 0;
        for (int py = 0; py < s; ++py)
            for (int px = 0; px < s; ++px) {
                const int hv = (d->Hout - py + s - 1) / s, wv = (d->Wout - px + s - 1) / s;
                if (d->Hout <= py || d->Wout <= px) continue;
                if (np >= MAX_PHASES) return 3;
                p.ph_tap0[np] = nt;
                p.ph_hv[np] = (short)hv;
                p.ph_wv[np] = (short)wv;
                p.ph_ooy[np] = (short)py;
                p.ph_oox[np] = (short)px;
                if (hv > dom_h) dom_h = hv;
                if (wv > dom_w) dom_w = wv;
                for (int ky = 0; ky < d->KH; ++ky) {
                    const int ny = py + d->pad - ky * d->dil;
                    if (pos_mod(ny, s) != 0) continue;
                    for (int kx = 0; kx < d->KW; ++kx) {
                        const int nx = px + d->pad - kx * d->dil;
                        if (pos_mod(nx, s) != 0) continue;
                        if (nt >= MAX_TAPS) return 3;
                        p.tap_dy[nt] = (short)floor_div(ny, s);
                        p.tap_dx[nt] = (short)floor_div(nx, s);
                        p.tap_w[nt] = (short)(ky * d->KW + kx);
                        p.tap_map[nt] = 0;
                        ++nt;
                    }
                }
                ++np;
            }
        p.nphases = np;
        p.ph_tap0[np] = nt;
        p.per_tap_map = 0;
    }
    if (p.nphases < 1) return 3;
    pick_tile(dom_h, dom_w, &p.TH, &p.TW);
    // ---- patch mode (see ConvTcParams::patch): 3x3, stride 1, dilation 1, one source of at most 64 channels, one N tile, all nine
    // B tiles resident beside at least three patches, enough pixel tiles per CTA (the plan_ring rule) ----
    {
        static const bool allow = env_on("CNB_TC_PATCH") && env_on("CNB_B_RESIDENT");
        const int b_bytes = BN * BK * 2;
        bool ok = allow && s == 1 && d->dil == 1 && d->KH == 3 && d->KW == 3 && d->nsrc == 1 && d->src_c[0] <= BK && d->N <= BN &&
                  p.nphases == 1 && p.ph_tap0[1] == 9;
        if (ok) {
            // a tile at least two rows high whose rows are whole swizzle atoms (TW >= 8): prefer TH = 8 (patch overhead 10 / 8)
            long best = -1;
            int bth = 0, btw = 0;
            for (int tw = 64; tw >= 8; tw >>= 1) {
                const int th = BM / tw;
                const long cover = (long)cnb_div_up(dom_w, tw) * tw * cnb_div_up(dom_h, th) * (th + 2);
                if (best < 0 || cover < best) best = cover, bth = th, btw = tw;
            }
            const int patch_bytes = (bth + 2) * btw * 128;
            const long tiles = (long)d->B * cnb_div_up(dom_h, bth) * cnb_div_up(dom_w, btw);
            ok = 9L * b_bytes + 3L * patch_bytes <= (long)(BN == 256 ? Cfg<256>::RING_BYTES : BN == 128 ? Cfg<128>::RING_BYTES
                                                            : BN == 64 ? Cfg<64>::RING_BYTES : Cfg<32>::RING_BYTES) &&
                 tiles >= 4L * num_sms();
            if (ok) {
                p.TH = bth, p.TW = btw;
                p.patch = 1;
                p.patch_bytes = patch_bytes;
                int miny = 0;
                for (int t = 0; t < 9; ++t) miny = p.tap_dy[t] < miny ? p.tap_dy[t] : miny;
                p.patch_y = (short)miny;
                int ng = 0;
                for (int gq = 0; gq < 3; ++gq)
                    for (int m = 0; m < 3; ++m) p.patch_tap[gq][m] = -1;
                for (int t = 0; t < 9; ++t) {
                    int gq = 0;
                    while (gq < ng && p.patch_dx[gq] != p.tap_dx[t]) ++gq;
                    if (gq == ng) {
                        if (ng == 3) return 3;
                        p.patch_dx[ng++] = p.tap_dx[t];
                    }
                    int m = 0;
                    while (m < 3 && p.patch_tap[gq][m] >= 0) ++m;
                    if (m == 3 || p.tap_dy[t] - miny > 2) return 3;
                    p.patch_tap[gq][m] = (short)t;
                    p.patch_row[gq][m] = (short)(p.tap_dy[t] - miny);
                }
                if (ng != 3) return 3;
            }
        }
    }
    int mt = 0;
    for (int ph = 0; ph < p.nphases; ++ph) {
        p.ph_tile0[ph] = mt;
        p.ph_tiles_h[ph] = (short)cnb_div_up(p.ph_hv[ph], p.TH);
        p.ph_tiles_w[ph] = (short)cnb_div_up(p.ph_wv[ph], p.TW);
        mt += d->B * p.ph_tiles_h[ph] * p.ph_tiles_w[ph];
    }
    p.ph_tile0[p.nphases] = mt;
    p.m_tiles = mt;

    // ---- operand maps ----
    int koff = 0;
    for (int si = 0; si < d->nsrc; ++si) {
        p.chunks[si] = cnb_div_up(d->src_c[si], BK);
        p.koff[si] = koff;
        koff += d->src_c[si];
    }
    p.nsrc = d->nsrc;
    if (p.per_tap_map) {
        for (int t = 0; t < p.ph_tap0[1]; ++t) {
            const int ky = p.tap_w[t] / d->KW, kx = p.tap_w[t] % d->KW;
            const int ry = pos_mod(ky * d->dil - d->pad, s), rx = pos_mod(kx * d->dil - d->pad, s);
            if (make_subgrid_map(&p.tmA[t], d->src[0], d->src_c[0], d->Win, d->Hin, d->B, d->src_stride[0], s, ry, rx, p.TW, p.TH)) return 2;
        }
    } else {
        for (int si = 0; si < d->nsrc; ++si)
            if (make_act_map(&p.tmA[si], d->src[si], d->src_c[si], d->Win, d->Hin, d->B, d->src_stride[si], p.TW, p.TH)) return 2;
    }
    if (p.patch && make_act_map(&p.tmA[1], d->src[0], d->src_c[0], d->Win, d->Hin, d->B, d->src_stride[0], p.TW, p.TH + 2)) return 2;
    if (make_weight_map(&p.tmB, d->w_packed, koff, d->N, taps, d->w_row_stride, d->w_tap_stride, BN)) return 2;

    p.N = d->N;
    p.Hout = d->Hout;
    p.Wout = d->Wout;
    p.out = reinterpret_cast<bf16_t*>(d->out);
    p.out_stride = d->out_stride;
    p.bias = d->bias;
    p.ep_scale = d->ep_scale, p.ep_shift = d->ep_shift, p.ep_act = d->ep_act;
    if (d->ep_scale && (!d->ep_shift || d->bias || d->nout > 0 || d->stats)) return 3;
    p.vec_ok = (d->out_stride % 8 == 0 && reinterpret_cast<uintptr_t>(d->out) % 16 == 0) ? 1 : 0;
    p.nseg = d->nout;
    if (d->nout > 0) {
        int c0 = 0;
        for (int i = 0; i < d->nout; ++i) {
            p.seg_begin[i] = c0;
            p.seg_out[i] = reinterpret_cast<bf16_t*>(d->out_seg[i]);
            p.seg_stride[i] = d->out_seg_stride[i];
            c0 += d->out_seg_c[i];
        }
        p.seg_begin[d->nout] = c0;
        p.vec_ok = 1;  // eligible(): every segment has a 16-byte pitch and base
        p.out = p.seg_out[0];
        p.out_stride = p.seg_stride[0];
    }
    p.log2_tw = 0;
    while ((1 << p.log2_tw) < p.TW) ++p.log2_tw;
    p.stats = nullptr;
    if (d->stats) {
        if (!stats_ok(d) || !p.vec_ok) return 3;
        p.stats = reinterpret_cast<float*>(d->stats);
    }
    p.num_tiles = cnb_div_up(d->N, BN) * p.m_tiles;
    {
        static const bool staged = env_on("CNB_EPI_STAGED"), parsmem = env_on("CNB_EPI_PARSMEM");
        static const int dbg = [] {
            const char* e = getenv("CNB_EPI_DEBUG");
            return e ? atoi(e) : 0;
        }();
        p.dbg = dbg;
        static const int pf = [] {
            const char* e = getenv("CNB_TC_PREFETCH");
            return e ? atoi(e) : 0;  // measured slower (a64 +17 %, 960 -> 256 1x1 +38 %): off unless asked for
        }();
        p.prefetch = pf;
        p.staged = (staged && p.vec_ok) ? 1 : 0;
        p.par_smem = (parsmem && !p.stats && (p.bias || p.ep_scale) && d->N <= STATS_MAX_N) ? 1 : 0;
    }
    switch (BN) {
        case 256: return launch_bn<256>(p, stream);
        case 128: return launch_bn<128>(p, stream);
        case 64: return launch_bn<64>(p, stream);
        default: return launch_bn<32>(p, stream);
    }
}

}  // namespace tc
}  // namespace cnb
#endif  // CNB_EMU
