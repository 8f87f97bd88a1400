// Bulk-copy streaming versions of the BatchNorm kernels (sm_100a only, bf16 pixel-major [P][L] with L/8 a power of two <= 256).
//
// ncu on the 128-bit register kernels of k_vec.cuh (profiles/r01_ncu_full_summary.json): 80-112 registers, 23-33 % occupancy,
// 2.7-3.9 TB/s -- the bytes a thread can keep in flight are bounded by its registers.  Here the bytes in flight live in shared
// memory instead: one elected thread streams 16 KB tiles of each operand through a ring with 1-D bulk async copies
// (cp.async.bulk.shared::cluster.global, completion on an mbarrier), all 256 threads wait on the tile's barrier, pull their
// four 16-byte vectors out of shared memory (conflict-free), release the stage with one __syncthreads (the elected thread
// refills it at once) and only then do the arithmetic and the global stores.  Two CTAs per SM keep up to ~190 KB in flight
// per SM, against the ~40 KB that 6.5 TB/s x ~1 us needs.
//
// A thread's vectors of one tile are tid + 256*j, and tiles start at multiples of 1024 vectors, so with CV = L/8 dividing 256
// the thread stays on ONE channel group for its whole life: per-channel constants and partial sums live in registers, as in
// the k_vec.cuh kernels whose arithmetic these reproduce exactly.
#pragma once
#ifndef CNB_EMU
#include "cnb_common.cuh"

namespace cnb {
namespace st {

constexpr int THREADS = 256;
constexpr int VPT = 4;                    // vectors per thread per tile
constexpr int TILE_V = THREADS * VPT;     // 1024 x 16 B
constexpr int TILE_BYTES = TILE_V * 16;   // 16 KB per operand per stage

__device__ __forceinline__ uint32_t s_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void s_mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void s_mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void s_mbar_wait(uint32_t bar, uint32_t parity) {
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
        if (clock64() - t0 > 4000000000LL) __trap();  // a protocol bug must fail the launch, not hang the GPU
    }
}
__device__ __forceinline__ void s_bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes),
                 "r"(bar)
                 : "memory");
}

__device__ __forceinline__ void s_unpack(const uint4& r, float* v) {
    v[0] = cnb_bits2f(r.x << 16), v[1] = cnb_bits2f(r.x & 0xffff0000u);
    v[2] = cnb_bits2f(r.y << 16), v[3] = cnb_bits2f(r.y & 0xffff0000u);
    v[4] = cnb_bits2f(r.z << 16), v[5] = cnb_bits2f(r.z & 0xffff0000u);
    v[6] = cnb_bits2f(r.w << 16), v[7] = cnb_bits2f(r.w & 0xffff0000u);
}
__device__ __forceinline__ uint4 s_pack(const float* v) {
    return make_uint4(cnb_pack_bf16x2(v[0], v[1]), cnb_pack_bf16x2(v[2], v[3]), cnb_pack_bf16x2(v[4], v[5]), cnb_pack_bf16x2(v[6], v[7]));
}

// Streams NT operands (same flat length, total_v 16-byte vectors) through an S-stage ring.  body(gv0, r, n): gv0 = global vector
// index of this thread's first vector of the tile, r[t][j] = vector tid + 256*j of operand t, n = how many j are inside the tensor.
template <int NT, int S, typename F>
__device__ __forceinline__ void stream_tiles(const void* const (&src)[NT], long total_v, unsigned char* smem, F&& body) {
    __shared__ __align__(8) uint64_t full[S];
    const int tid = threadIdx.x;
    const long ntiles = (total_v + TILE_V - 1) / TILE_V;
    const long stride = gridDim.x;
    const uint32_t sm0 = s_u32(smem);
    auto issue = [&](long tile, int stage) {
        const long v0 = tile * TILE_V;
        const long left = total_v - v0;
        const uint32_t bytes = (uint32_t)(left < TILE_V ? left : TILE_V) * 16u;
        const uint32_t bar = s_u32(&full[stage]);
        s_mbar_expect_tx(bar, bytes * NT);
#pragma unroll
        for (int t = 0; t < NT; ++t)
            s_bulk_g2s(sm0 + (uint32_t)((stage * NT + t) * TILE_BYTES), reinterpret_cast<const unsigned char*>(src[t]) + v0 * 16, bytes, bar);
    };
    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < S; ++s) s_mbar_init(s_u32(&full[s]), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < S; ++s) {
            const long tile = blockIdx.x + s * stride;
            if (tile < ntiles) issue(tile, s);
        }
    }
    int stage = 0;
    uint32_t phase = 0;
    for (long tile = blockIdx.x; tile < ntiles; tile += stride) {
        s_mbar_wait(s_u32(&full[stage]), phase);
        const long v0 = tile * TILE_V + tid;
        uint4 r[NT][VPT];
        int n = 0;
#pragma unroll
        for (int j = 0; j < VPT; ++j)
            if (v0 + j * THREADS < total_v) {
                n = j + 1;
#pragma unroll
                for (int t = 0; t < NT; ++t)
                    r[t][j] = *reinterpret_cast<const uint4*>(smem + (stage * NT + t) * TILE_BYTES + (tid + j * THREADS) * 16);
            }
        __syncthreads();  // every thread holds its vectors in registers: the stage is free again
        if (tid == 0) {
            const long next = tile + S * stride;
            if (next < ntiles) issue(next, stage);
        }
        body(v0, r, n);
        if (++stage == S) {
            stage = 0;
            phase ^= 1u;
        }
    }
}

template <int NT, int S>
__host__ __device__ constexpr int smem_bytes() {
    return NT * S * TILE_BYTES;
}
constexpr int S1 = 5, S2 = 3;  // stages for one / two streamed operands (80 KB / 96 KB per CTA, two CTAs per SM)

// per-thread channel table: column c0 + j belongs to channel (c0 + j) / ch_div; columns past C*ch_div are row padding (-1)
__device__ __forceinline__ void s_channels(int CV, int C, int ch_div, int (&chn)[8]) {
    const int c0 = (int)(threadIdx.x % CV) * 8;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int ch = (c0 + j) / ch_div;
        chn[j] = ch < C ? ch : -1;
    }
}

__device__ __forceinline__ void s_flush_sums(float* sh, int C, const int (&chn)[8], const float (&a)[8], const float (&b)[8], float* out) {
    __syncthreads();
#pragma unroll
    for (int j = 0; j < 8; ++j)
        if (chn[j] >= 0) {
            atomicAdd(&sh[chn[j]], a[j]);
            atomicAdd(&sh[C + chn[j]], b[j]);
        }
    __syncthreads();
    for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) atomicAdd(&out[i], sh[i]);
}

// sums[0..C) += sum x, sums[C..2C) += sum x^2
__global__ void __launch_bounds__(THREADS, 2) bn_stats_stream_kernel(const bf16_t* __restrict__ x, long total_v, int CV, int C, int ch_div,
                                                                    float* __restrict__ sums) {
    CNB_PDL_SYNC();
    CNB_DYN_SMEM(smem);
    float* sh = reinterpret_cast<float*>(smem + smem_bytes<1, S1>());
    for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) sh[i] = 0.f;
    int chn[8];
    s_channels(CV, C, ch_div, chn);
    float s[8], q[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) s[j] = 0.f, q[j] = 0.f;
    const void* const src[1] = {x};
    stream_tiles<1, S1>(src, total_v, smem, [&](long, uint4(&r)[1][VPT], int n) {
#pragma unroll
        for (int u = 0; u < VPT; ++u)
            if (u < n) {
                float v[8];
                s_unpack(r[0][u], v);
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    s[j] += v[j];
                    q[j] = fmaf(v[j], v[j], q[j]);
                }
            }
    });
    s_flush_sums(sh, C, chn, s, q, sums);
}

// y = act(x*scale + shift) (+ residual)
template <bool RES>
__global__ void __launch_bounds__(THREADS, 2) bn_act_fwd_stream_kernel(const bf16_t* __restrict__ x, const float* __restrict__ scale,
                                                                      const float* __restrict__ shift, const bf16_t* __restrict__ residual,
                                                                      bf16_t* __restrict__ y, long total_v, int CV, int C, int ch_div,
                                                                      int act) {
    CNB_PDL_SYNC();
    CNB_DYN_SMEM(smem);
    int chn[8];
    s_channels(CV, C, ch_div, chn);
    float sc[8], sf[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        sc[j] = chn[j] >= 0 ? scale[chn[j]] : 0.f;  // padding columns come out as act(0) = 0
        sf[j] = chn[j] >= 0 ? shift[chn[j]] : 0.f;
    }
    auto body = [&](long v0, auto& r, int n) {
        constexpr int LAST = RES ? 1 : 0;
#pragma unroll
        for (int u = 0; u < VPT; ++u)
            if (u < n) {
                float v[8], w[8];
                s_unpack(r[0][u], v);
                if (RES) s_unpack(r[LAST][u], w);
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    float z = fmaf(v[j], sc[j], sf[j]);
                    z = cnb_act_t<bf16_t>(z, act);
                    v[j] = RES ? z + w[j] : z;
                }
                *reinterpret_cast<uint4*>(y + (v0 + (long)u * THREADS) * 8) = s_pack(v);
            }
    };
    if constexpr (RES) {
        const void* const src[2] = {x, residual};
        stream_tiles<2, S2>(src, total_v, smem, body);
    } else {
        const void* const src[1] = {x};
        stream_tiles<1, S1>(src, total_v, smem, body);
    }
}

// Training-mode forward with the statistics finalisation folded in (one launch instead of bn_finalize + apply): every thread
// derives scale/shift of ITS eight channels from the batch sums; CTA 0 also publishes mean / rstd / scale / shift for the backward
// pass and updates the running statistics (it is the only CTA that touches them).
template <bool RES>
__global__ void __launch_bounds__(THREADS, 2) bn_train_fwd_stream_kernel(const bf16_t* __restrict__ x, const float* __restrict__ sums,
                                                                        long count, const float* __restrict__ gamma,
                                                                        const float* __restrict__ beta, float eps, float momentum,
                                                                        float* running_mean, float* running_var, float* __restrict__ save_mean,
                                                                        float* __restrict__ save_rstd, float* __restrict__ scale,
                                                                        float* __restrict__ shift, const bf16_t* __restrict__ residual,
                                                                        bf16_t* __restrict__ y, long total_v, int CV, int C, int ch_div,
                                                                        int act) {
    CNB_PDL_SYNC();
    CNB_DYN_SMEM(smem);
    const float inv = 1.0f / (float)count;
    if (blockIdx.x == 0) {
        for (int c = threadIdx.x; c < C; c += blockDim.x) {
            const float mean = sums[c] * inv;
            const float var = fmaxf(sums[C + c] * inv - mean * mean, 0.f);
            if (running_mean) {
                const float unbiased = count > 1 ? var * ((float)count / (float)(count - 1)) : var;
                running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * mean;
                running_var[c] = (1.f - momentum) * running_var[c] + momentum * unbiased;
            }
            const float rstd = rsqrtf(var + eps);
            const float g = gamma ? gamma[c] : 1.f, b = beta ? beta[c] : 0.f;
            save_mean[c] = mean;
            save_rstd[c] = rstd;
            scale[c] = g * rstd;
            shift[c] = b - mean * g * rstd;
        }
    }
    int chn[8];
    s_channels(CV, C, ch_div, chn);
    float sc[8], sf[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        if (chn[j] >= 0) {
            const int c = chn[j];
            const float mean = sums[c] * inv;
            const float var = fmaxf(sums[C + c] * inv - mean * mean, 0.f);
            const float rstd = rsqrtf(var + eps);
            const float g = gamma ? gamma[c] : 1.f, b = beta ? beta[c] : 0.f;
            sc[j] = g * rstd;
            sf[j] = b - mean * g * rstd;
        } else {
            sc[j] = 0.f, sf[j] = 0.f;
        }
    }
    auto body = [&](long v0, auto& r, int n) {
        constexpr int LAST = RES ? 1 : 0;
#pragma unroll
        for (int u = 0; u < VPT; ++u)
            if (u < n) {
                float v[8], w[8];
                s_unpack(r[0][u], v);
                if (RES) s_unpack(r[LAST][u], w);
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    float z = fmaf(v[j], sc[j], sf[j]);
                    z = cnb_act_t<bf16_t>(z, act);
                    v[j] = RES ? z + w[j] : z;
                }
                *reinterpret_cast<uint4*>(y + (v0 + (long)u * THREADS) * 8) = s_pack(v);
            }
    };
    if constexpr (RES) {
        const void* const src[2] = {x, residual};
        stream_tiles<2, S2>(src, total_v, smem, body);
    } else {
        const void* const src[1] = {x};
        stream_tiles<1, S1>(src, total_v, smem, body);
    }
}

// dsums[0..C) += sum dz, dsums[C..2C) += sum dz*xhat, dz = dy * act'(x*A + Bc)
__global__ void __launch_bounds__(THREADS, 2) bn_act_bwd_reduce_stream_kernel(const bf16_t* __restrict__ x, const bf16_t* __restrict__ dy,
                                                                             const float* __restrict__ mean, const float* __restrict__ rstd,
                                                                             const float* __restrict__ gamma, const float* __restrict__ beta,
                                                                             long total_v, int CV, int C, int ch_div, int act,
                                                                             float* __restrict__ dsums) {
    CNB_PDL_SYNC();
    CNB_DYN_SMEM(smem);
    float* sh = reinterpret_cast<float*>(smem + smem_bytes<2, S2>());
    for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) sh[i] = 0.f;
    int chn[8];
    s_channels(CV, C, ch_div, chn);
    float mu[8], A[8], Bc[8], s[8], sx[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int cc = chn[j] >= 0 ? chn[j] : 0;
        const float gj = gamma ? gamma[cc] : 1.f, bj = beta ? beta[cc] : 0.f;
        mu[j] = mean[cc];
        A[j] = gj * rstd[cc];
        Bc[j] = bj - mu[j] * A[j];
        s[j] = 0.f, sx[j] = 0.f;
    }
    const void* const src[2] = {x, dy};
    stream_tiles<2, S2>(src, total_v, smem, [&](long, uint4(&r)[2][VPT], int n) {
#pragma unroll
        for (int u = 0; u < VPT; ++u)
            if (u < n) {
                float xv[8], dv[8];
                s_unpack(r[0][u], xv);
                s_unpack(r[1][u], dv);
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    float dz = dv[j];
                    if (act) dz *= cnb_act_grad_t<bf16_t>(fmaf(xv[j], A[j], Bc[j]), act);
                    s[j] += dz;
                    sx[j] = fmaf(dz, xv[j] - mu[j], sx[j]);
                }
            }
    });
#pragma unroll
    for (int j = 0; j < 8; ++j) sx[j] *= rstd[chn[j] >= 0 ? chn[j] : 0];
    s_flush_sums(sh, C, chn, s, sx, dsums);
}

// dx = A*dz - x*K1 + Q  (see bn_act_bwd_apply_vec_kernel)
__global__ void __launch_bounds__(THREADS, 2) bn_act_bwd_apply_stream_kernel(const bf16_t* __restrict__ x, const bf16_t* __restrict__ dy,
                                                                            const float* __restrict__ mean, const float* __restrict__ rstd,
                                                                            const float* __restrict__ gamma, const float* __restrict__ beta,
                                                                            const float* __restrict__ dsums, float inv_count,
                                                                            bf16_t* __restrict__ dx, long total_v, int CV, int C, int ch_div,
                                                                            int act, int train_stats) {
    CNB_PDL_SYNC();
    CNB_DYN_SMEM(smem);
    int chn[8];
    s_channels(CV, C, ch_div, chn);
    float A[8], Bc[8], K1[8], Q[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int ch = chn[j];
        if (ch >= 0) {
            const float gj = gamma ? gamma[ch] : 1.f, bj = beta ? beta[ch] : 0.f;
            const float m = mean[ch], rs = rstd[ch];
            A[j] = gj * rs;
            Bc[j] = bj - m * A[j];
            K1[j] = train_stats ? A[j] * rs * dsums[C + ch] * inv_count : 0.f;
            Q[j] = train_stats ? m * K1[j] - A[j] * dsums[ch] * inv_count : 0.f;
        } else {  // row padding: gradient zero
            A[j] = 0.f, Bc[j] = 0.f, K1[j] = 0.f, Q[j] = 0.f;
        }
    }
    const void* const src[2] = {x, dy};
    stream_tiles<2, S2>(src, total_v, smem, [&](long v0, uint4(&r)[2][VPT], int n) {
#pragma unroll
        for (int u = 0; u < VPT; ++u)
            if (u < n) {
                float xv[8], dv[8];
                s_unpack(r[0][u], xv);
                s_unpack(r[1][u], dv);
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    float dz = dv[j];
                    if (act) dz *= cnb_act_grad_t<bf16_t>(fmaf(xv[j], A[j], Bc[j]), act);
                    xv[j] = fmaf(A[j], dz, fmaf(-xv[j], K1[j], Q[j]));
                }
                *reinterpret_cast<uint4*>(dx + (v0 + (long)u * THREADS) * 8) = s_pack(xv);
            }
    });
}

// is [P][L] bf16 a shape these kernels take?  (C <= 2048 keeps the per-CTA channel sums next to the ring)
static inline bool eligible(int L, int C, int ch_div, int dtype, long total_v) {
    if (dtype != CNB_BF16 || L % 8 != 0) return false;
    const int CV = L / 8;
    if (CV > THREADS || THREADS % CV != 0) return false;
    if ((long)C * ch_div > L || C > 2048) return false;
    return total_v >= 4L * TILE_V;  // tiny tensors: the register kernels launch fewer, fuller CTAs
}
static inline int grid(long total_v) {
    const long ntiles = (total_v + TILE_V - 1) / TILE_V;
    const long cap = 2L * CNB_NUM_SMS;
    return (int)(ntiles < cap ? ntiles : cap);
}
// the reducing kernels end with 2C shared + 2C global atomics per CTA: on a small tensor (level c: 3.5 tiles per CTA at the grid above)
// that flush was most of the launch (28.6 us for 33 MB against 20.8 us for the apply kernel that moves half as much again), so they
// run at least 8 tiles per CTA
static inline int grid_reduce(long total_v) {
    const long ntiles = (total_v + TILE_V - 1) / TILE_V;
    const long cap = 2L * CNB_NUM_SMS;
    long g = (ntiles + 7) / 8;
    if (g > cap) g = cap;
    return (int)(g < 1 ? 1 : g);
}

}  // namespace st
}  // namespace cnb
#endif  // CNB_EMU
