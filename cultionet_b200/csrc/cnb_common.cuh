// Shared device/host helpers for the cultionet_b200 kernels (sm_100a).
//
// Every kernel in this directory is written against NHWC ("pixel-major") activations: a tensor is
// [P = B*H*W pixels][C channels] with the channel index contiguous, in either fp32 (parity mode) or
// bf16 (throughput mode); statistics, accumulators and parameter gradients are always fp32.
#pragma once

#ifndef CNB_EMU
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
// Every kernel of this library is launched with programmatic stream serialisation (programmatic dependent launch): its CTAs may be
// scheduled while the previous kernel of the stream is still draining, run their prologue, and block in CNB_PDL_SYNC() -- the first
// statement that may touch global memory -- until the previous grid has completed and its writes are visible.  A step is ~1200
// mostly short dependent kernels (a 268 MB reduction takes 68 us of which ~20 us are ramp-up and tail); this overlaps the ramp of
// kernel i+1 with the tail of kernel i, also inside the captured CUDA graph (programmatic edges).  CNB_PDL=0 in the environment
// restores plain launches.
inline bool cnb_pdl_enabled() {
    static const bool on = [] {
        const char* e = getenv("CNB_PDL");
        return !(e && e[0] == '0');
    }();
    return on;
}
template <typename... KArgs, typename... Args>
inline void cnb_launch_kernel(void (*kfn)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args&&... args) {
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = cnb_pdl_enabled() ? 1 : 0;
    cudaLaunchKernelEx(&cfg, kfn, static_cast<KArgs>(args)...);
}
#define CNB_LAUNCH(kfn, grid, block, smem, stream, ...) (cnb_count_launch(), cnb_launch_kernel(kfn, grid, block, smem, stream, __VA_ARGS__))
// wait for the previous grid of the stream (no-op without a programmatic dependency), then allow the next grid to be scheduled
#define CNB_PDL_SYNC()                                              \
    do {                                                            \
        asm volatile("griddepcontrol.wait;" ::: "memory");          \
        asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); \
    } while (0)
#define CNB_DYN_SMEM(name) extern __shared__ __align__(1024) unsigned char name[]
#define CNB_MEMSET_ASYNC(ptr, val, bytes, stream) cudaMemsetAsync((ptr), (val), (bytes), (stream))
#define CNB_PEEK_ERROR() cudaPeekAtLastError()
#define CNB_CLEAR_ERROR() ((void)cudaGetLastError())
#define CNB_ERROR_STRING(e) cudaGetErrorString(e)
#endif

#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>

// number of kernels this library has launched (reported by bench.py as `gpu_launches`)
inline std::atomic<long long>& cnb_launch_counter() {
    static std::atomic<long long> n{0};
    return n;
}
inline void cnb_count_launch() { cnb_launch_counter().fetch_add(1, std::memory_order_relaxed); }

#include "../../include/cultionet_b200.h"

typedef __nv_bfloat16 bf16_t;

// ---------------------------------------------------------------------------------------------
// error reporting: every extern "C" entry point returns 0 or a CNB_ERR_* code and leaves a message
// ---------------------------------------------------------------------------------------------
inline char* cnb_err_buf() {
    static thread_local char buf[512] = {0};
    return buf;
}
#define CNB_FAIL(code, ...)                               \
    do {                                                  \
        snprintf(cnb_err_buf(), 512, __VA_ARGS__);        \
        return (code);                                    \
    } while (0)
#define CNB_REQUIRE(cond, ...)                            \
    do {                                                  \
        if (!(cond)) CNB_FAIL(CNB_ERR_INVALID, __VA_ARGS__); \
    } while (0)
#define CNB_CHECK_LAUNCH(name)                                                          \
    do {                                                                                \
        cudaError_t e__ = CNB_PEEK_ERROR();                                             \
        if (e__ != cudaSuccess) {                                                       \
            CNB_CLEAR_ERROR();                                                          \
            CNB_FAIL(CNB_ERR_CUDA, "%s: %s", name, CNB_ERROR_STRING(e__));              \
        }                                                                               \
    } while (0)

// dispatch on the activation dtype enum
#define CNB_DISPATCH_DTYPE(dtype, ...)                                 \
    do {                                                               \
        if ((dtype) == CNB_F32) {                                      \
            typedef float T;                                           \
            __VA_ARGS__                                                \
        } else if ((dtype) == CNB_BF16) {                              \
            typedef bf16_t T;                                          \
            __VA_ARGS__                                                \
        } else {                                                       \
            CNB_FAIL(CNB_ERR_INVALID, "unsupported dtype %d", (int)(dtype)); \
        }                                                              \
    } while (0)

// ---------------------------------------------------------------------------------------------
// element access
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float cnb_ld(const float* p) { return *p; }
__device__ __forceinline__ float cnb_ld(const bf16_t* p) { return __bfloat162float(*p); }
__device__ __forceinline__ void cnb_st(float* p, float v) { *p = v; }
__device__ __forceinline__ void cnb_st(bf16_t* p, float v) { *p = __float2bfloat16(v); }
// value after a round trip through the storage type (what a later kernel will read back)
__device__ __forceinline__ float cnb_round(float v, const float*) { return v; }
__device__ __forceinline__ float cnb_round(float v, const bf16_t*) { return __bfloat162float(__float2bfloat16(v)); }

// ---------------------------------------------------------------------------------------------
// 16-byte vector access (8 x bf16 or 4 x fp32): what every bandwidth-bound kernel moves per thread per step
// ---------------------------------------------------------------------------------------------
template <typename T>
struct cnb_vec {
    static constexpr int N = 16 / (int)sizeof(T);
};
__device__ __forceinline__ float cnb_bits2f(uint32_t u) {
    float f;
    memcpy(&f, &u, 4);
    return f;
}
__device__ __forceinline__ void cnb_ldv(const float* p, float* v) {
    const float4 r = *reinterpret_cast<const float4*>(p);
    v[0] = r.x, v[1] = r.y, v[2] = r.z, v[3] = r.w;
}
__device__ __forceinline__ void cnb_ldv(const bf16_t* p, float* v) {
    const uint4 r = *reinterpret_cast<const uint4*>(p);
    v[0] = cnb_bits2f(r.x << 16), v[1] = cnb_bits2f(r.x & 0xffff0000u);
    v[2] = cnb_bits2f(r.y << 16), v[3] = cnb_bits2f(r.y & 0xffff0000u);
    v[4] = cnb_bits2f(r.z << 16), v[5] = cnb_bits2f(r.z & 0xffff0000u);
    v[6] = cnb_bits2f(r.w << 16), v[7] = cnb_bits2f(r.w & 0xffff0000u);
}
// the same load split in two: the raw 16 bytes (what a thread keeps in flight: 4 registers for 8 bf16 values) and their expansion
template <typename T>
__device__ __forceinline__ uint4 cnb_ldraw(const T* p) {
    return *reinterpret_cast<const uint4*>(p);
}
__device__ __forceinline__ void cnb_expand(const uint4& r, float* v, const float*) {
    v[0] = cnb_bits2f(r.x), v[1] = cnb_bits2f(r.y), v[2] = cnb_bits2f(r.z), v[3] = cnb_bits2f(r.w);
}
__device__ __forceinline__ void cnb_expand(const uint4& r, float* v, const bf16_t*) {
    v[0] = cnb_bits2f(r.x << 16), v[1] = cnb_bits2f(r.x & 0xffff0000u);
    v[2] = cnb_bits2f(r.y << 16), v[3] = cnb_bits2f(r.y & 0xffff0000u);
    v[4] = cnb_bits2f(r.z << 16), v[5] = cnb_bits2f(r.z & 0xffff0000u);
    v[6] = cnb_bits2f(r.w << 16), v[7] = cnb_bits2f(r.w & 0xffff0000u);
}
__device__ __forceinline__ void cnb_stv(float* p, const float* v) { *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]); }
__device__ __forceinline__ uint32_t cnb_pack_bf16x2(float lo, float hi) {
    const __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
    uint32_t u;
    memcpy(&u, &h, 4);
    return u;
}
__device__ __forceinline__ void cnb_stv(bf16_t* p, const float* v) {
    *reinterpret_cast<uint4*>(p) =
        make_uint4(cnb_pack_bf16x2(v[0], v[1]), cnb_pack_bf16x2(v[2], v[3]), cnb_pack_bf16x2(v[4], v[5]), cnb_pack_bf16x2(v[6], v[7]));
}
// 16-byte asynchronous global -> shared copy (LDGSTS): a thread can keep many of these in flight while it issues more
__device__ __forceinline__ void cnb_cp_async16(void* smem_dst, const void* gsrc) {
#ifdef CNB_EMU
    memcpy(smem_dst, gsrc, 16);
#else
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc) : "memory");
#endif
}
__device__ __forceinline__ void cnb_cp_async_wait_all() {
#ifndef CNB_EMU
    asm volatile("cp.async.wait_all;" ::: "memory");
#endif
}
__device__ __forceinline__ bool cnb_aligned16_dev(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }
static inline bool cnb_aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

__device__ __forceinline__ float cnb_exp(float x) {
#ifdef CNB_EMU
    return expf(x);
#else
    return __expf(x);
#endif
}
__device__ __forceinline__ float cnb_sigmoid(float x) { return 1.0f / (1.0f + cnb_exp(-x)); }
// One MUFU op instead of two (ex2 + rcp): sigmoid(x) = 0.5 + 0.5*tanh(x/2) with tanh.approx (abs. error < 3e-4, far inside a bf16 ulp).
// Only the bf16 (throughput-mode) kernels use it; the fp32 parity path keeps the exact form.
__device__ __forceinline__ float cnb_sigmoid_fast(float x) {
#ifdef CNB_EMU
    return 1.0f / (1.0f + expf(-x));
#else
    float t;
    asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(0.5f * x));
    return fmaf(0.5f, t, 0.5f);
#endif
}
template <typename T>
__device__ __forceinline__ float cnb_sigmoid_t(float x) {
    return sizeof(T) == 2 ? cnb_sigmoid_fast(x) : cnb_sigmoid(x);
}
template <typename T>
__device__ __forceinline__ float cnb_silu_t(float x) {
    return x * cnb_sigmoid_t<T>(x);
}
template <typename T>
__device__ __forceinline__ float cnb_silu_grad_t(float x) {
    const float s = cnb_sigmoid_t<T>(x);
    return s * (1.0f + x * (1.0f - s));
}
// ---------------------------------------------------------------------------------------------
// activation codes (cultionet_b200.h CNB_ACT_*): the reference builds `getattr(torch.nn, activation_type)()` (activations.py:5-24);
// SiLU is its default and the code every fused kernel is tuned for -- the other codes take one out-of-line call per element.
// ---------------------------------------------------------------------------------------------
#ifdef CNB_EMU
#define CNB_NOINLINE
#else
#define CNB_NOINLINE __noinline__
#endif
__device__ CNB_NOINLINE float cnb_act_other(float z, int act) {
    switch (act) {
        case CNB_ACT_RELU: return z > 0.f ? z : 0.f;
        case CNB_ACT_LEAKY_RELU: return z > 0.f ? z : 0.01f * z;
        case CNB_ACT_GELU: return 0.5f * z * (1.0f + erff(z * 0.70710678118654752f));
        case CNB_ACT_MISH: {
            const float sp = z > 20.f ? z : log1pf(expf(z));  // softplus, torch threshold 20
            return z * tanhf(sp);
        }
        case CNB_ACT_ELU: return z > 0.f ? z : expm1f(z);
        case CNB_ACT_TANH: return tanhf(z);
        case CNB_ACT_SIGMOID: return 1.0f / (1.0f + expf(-z));
        case CNB_ACT_HARDSWISH: return z <= -3.f ? 0.f : (z >= 3.f ? z : z * (z + 3.f) * (1.0f / 6.0f));
        default: return z;
    }
}
__device__ CNB_NOINLINE float cnb_act_grad_other(float z, int act) {
    switch (act) {
        case CNB_ACT_RELU: return z > 0.f ? 1.f : 0.f;
        case CNB_ACT_LEAKY_RELU: return z > 0.f ? 1.f : 0.01f;
        case CNB_ACT_GELU:
            return 0.5f * (1.0f + erff(z * 0.70710678118654752f)) + z * 0.3989422804014327f * expf(-0.5f * z * z);
        case CNB_ACT_MISH: {
            const float sp = z > 20.f ? z : log1pf(expf(z));
            const float t = tanhf(sp);
            const float sg = 1.0f / (1.0f + expf(-z));  // d softplus / dz
            return t + z * (1.0f - t * t) * sg;
        }
        case CNB_ACT_ELU: return z > 0.f ? 1.f : expf(z);
        case CNB_ACT_TANH: {
            const float t = tanhf(z);
            return 1.0f - t * t;
        }
        case CNB_ACT_SIGMOID: {
            const float sg = 1.0f / (1.0f + expf(-z));
            return sg * (1.0f - sg);
        }
        case CNB_ACT_HARDSWISH: return z <= -3.f ? 0.f : (z >= 3.f ? 1.f : (2.0f * z + 3.0f) * (1.0f / 6.0f));  // torch: open interval
        default: return 1.f;
    }
}
// act(z) / act'(z) for an activation code; T selects the SiLU flavour (bf16 kernels: one-MUFU sigmoid)
template <typename T>
__device__ __forceinline__ float cnb_act_t(float z, int act) {
    return act == CNB_ACT_SILU ? cnb_silu_t<T>(z) : (act == CNB_ACT_NONE ? z : cnb_act_other(z, act));
}
template <typename T>
__device__ __forceinline__ float cnb_act_grad_t(float z, int act) {
    return act == CNB_ACT_SILU ? cnb_silu_grad_t<T>(z) : (act == CNB_ACT_NONE ? 1.f : cnb_act_grad_other(z, act));
}
__device__ __forceinline__ float cnb_silu(float x) { return x * cnb_sigmoid(x); }
// d/dx [x * sigmoid(x)]
__device__ __forceinline__ float cnb_silu_grad(float x) {
    float s = cnb_sigmoid(x);
    return s * (1.0f + x * (1.0f - s));
}

// ---------------------------------------------------------------------------------------------
// Mixed-precision FMA (sm_100: fma.rn.f32.bf16 -> SASS FHFMA.BF16): fp32 accumulator, bf16 operands read as the halves of packed
// 32-bit words.  A bf16 operand is never unpacked: the (shift, mask) pair per element of the usual bf16 -> fp32 path runs on the
// half-rate integer pipe and was half of the instructions of the neighbourhood-attention kernels (ncu: 70-75 % issue-slot busy).
// ---------------------------------------------------------------------------------------------
// acc + a.lo * b.lo + a.hi * b.hi
__device__ __forceinline__ float cnb_fma2_bf16(uint32_t a, uint32_t b, float acc) {
#ifdef CNB_EMU
    acc = fmaf(cnb_bits2f(a << 16), cnb_bits2f(b << 16), acc);
    return fmaf(cnb_bits2f(a & 0xffff0000u), cnb_bits2f(b & 0xffff0000u), acc);
#else
    asm("{\n\t.reg .b16 al, ah, bl, bh;\n\tmov.b32 {al, ah}, %1;\n\tmov.b32 {bl, bh}, %2;\n\t"
        "fma.rn.f32.bf16 %0, al, bl, %0;\n\tfma.rn.f32.bf16 %0, ah, bh, %0;\n\t}"
        : "+f"(acc)
        : "r"(a), "r"(b));
    return acc;
#endif
}
// (acc_lo, acc_hi) += w * (b.lo, b.hi) with w = the LOW (HI = false) or HIGH (HI = true) half of wp
template <bool HI>
__device__ __forceinline__ void cnb_axpy2_bf16(uint32_t wp, uint32_t b, float& acc_lo, float& acc_hi) {
#ifdef CNB_EMU
    const float w = HI ? cnb_bits2f(wp & 0xffff0000u) : cnb_bits2f(wp << 16);
    acc_lo = fmaf(w, cnb_bits2f(b << 16), acc_lo);
    acc_hi = fmaf(w, cnb_bits2f(b & 0xffff0000u), acc_hi);
#else
    if (HI)
        asm("{\n\t.reg .b16 wl, wh, bl, bh;\n\tmov.b32 {wl, wh}, %2;\n\tmov.b32 {bl, bh}, %3;\n\t"
            "fma.rn.f32.bf16 %0, wh, bl, %0;\n\tfma.rn.f32.bf16 %1, wh, bh, %1;\n\t}"
            : "+f"(acc_lo), "+f"(acc_hi)
            : "r"(wp), "r"(b));
    else
        asm("{\n\t.reg .b16 wl, wh, bl, bh;\n\tmov.b32 {wl, wh}, %2;\n\tmov.b32 {bl, bh}, %3;\n\t"
            "fma.rn.f32.bf16 %0, wl, bl, %0;\n\tfma.rn.f32.bf16 %1, wl, bh, %1;\n\t}"
            : "+f"(acc_lo), "+f"(acc_hi)
            : "r"(wp), "r"(b));
#endif
}

__device__ __forceinline__ float cnb_warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double cnb_warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float cnb_warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// counter-based random bits for the dropout kernels (k_pool_attn.cuh, k_na.cuh): state = device int64[2] {seed, step counter}
__device__ __forceinline__ uint64_t cnb_mix64(uint64_t z) {
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL;
    return z ^ (z >> 31);
}
__device__ __forceinline__ uint64_t cnb_rng_key(const int64_t* state, int site) {
    return cnb_mix64((uint64_t)state[0] ^ cnb_mix64((uint64_t)state[1] * 0x9e3779b97f4a7c15ULL + (uint64_t)(uint32_t)site));
}
// 16 random bits of element `e` of the stream `key`
__device__ __forceinline__ uint32_t cnb_rng_bits16(uint64_t key, uint64_t e) {
    const uint64_t h = cnb_mix64(key + (e >> 2) * 0x9e3779b97f4a7c15ULL);
    return (uint32_t)(h >> (16 * (e & 3))) & 0xffffu;
}

static inline int cnb_div_up(long a, long b) { return (int)((a + b - 1) / b); }
static inline int cnb_clamp_grid(long blocks, long cap) { return (int)(blocks < 1 ? 1 : (blocks > cap ? cap : blocks)); }

// 148 SMs on a B200; grids of grid-stride kernels are sized in multiples of this.
#define CNB_NUM_SMS 148
