// TEST INFRASTRUCTURE ONLY -- a CPU interpreter for the SIMT subset of CUDA the cultionet_b200
// kernels use (threadIdx/blockIdx, __syncthreads, warp shuffles, atomics, __shared__, bf16).
//
// tests/emu/build_emu.py force-includes this header and compiles cultionet_b200/csrc/cnb_api.cu as
// plain C++ into tests/emu/libcnb_emu.so.  The `-m "not gpu"` tests load that library explicitly to
// check kernel arithmetic and the Python wiring on a box without a GPU.  The product package never
// loads it (cultionet_b200/_lib.py only loads the nvcc-built library and raises when it is missing)
// and the tcgen05/TMA kernels are compiled out (#ifndef CNB_EMU): those are tested on the GPU only.
//
// Execution model: the blocks of a launch are distributed over a few OS threads; inside one block every
// CUDA thread is a fiber (hand-rolled x86-64 context switch) scheduled round-robin on that OS thread, so
// __syncthreads()/shuffles are cooperative yields and `__shared__` maps to `static thread_local`.
#pragma once
#ifndef CNB_EMU
#error "cnb_emu.h is only for -DCNB_EMU host builds"
#endif
#if !defined(__x86_64__)
#error "the fiber switch is written for x86-64"
#endif

#include <cuda_bf16.h>
#include <vector_functions.h>
#include <vector_types.h>

#include <atomic>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <thread>
#include <vector>

#undef __global__
#undef __device__
#undef __host__
#undef __shared__
#undef __forceinline__
#undef __launch_bounds__
#undef __restrict__
#undef __align__
#define __global__
#define __device__
#define __host__
#define __shared__ static thread_local
#define __forceinline__ inline __attribute__((always_inline))
#define __launch_bounds__(...)
#define __restrict__ __restrict
#define __align__(n) __attribute__((aligned(n)))
#undef __grid_constant__
#define __grid_constant__

#define CNB_MEMSET_ASYNC(ptr, val, bytes, stream) memset((ptr), (val), (bytes))
#define CNB_PDL_SYNC() ((void)0)
#define CNB_PEEK_ERROR() cudaSuccess
#define CNB_CLEAR_ERROR() ((void)0)
#define CNB_ERROR_STRING(e) "emu"

extern "C" void cnb_emu_switch(void** from_sp, void* to_sp);
asm(R"(
.text
.globl cnb_emu_switch
.type cnb_emu_switch,@function
cnb_emu_switch:
    pushq %rbp
    pushq %rbx
    pushq %r12
    pushq %r13
    pushq %r14
    pushq %r15
    movq %rsp, (%rdi)
    movq %rsi, %rsp
    popq %r15
    popq %r14
    popq %r13
    popq %r12
    popq %rbx
    popq %rbp
    ret
.size cnb_emu_switch, .-cnb_emu_switch
)");

namespace cnb_emu {

struct Warp {
    int live = 0, count = 0, gen = 0;
    uint64_t slot[32];
};

struct Block {
    uint3 bidx;
    dim3 bdim, gdim;
    int live = 0, bar_count = 0, bar_gen = 0;
    std::vector<Warp> warps;
    unsigned char* dyn_smem = nullptr;
};

struct Fiber {
    void* sp = nullptr;
    bool done = false;
    uint3 tidx;
    int warp = 0, lane = 0;
};

inline thread_local Block* g_blk = nullptr;
inline thread_local Fiber* g_cur = nullptr;
inline thread_local void* g_sched_sp = nullptr;
inline thread_local const std::function<void()>* g_fn = nullptr;

inline void yield() { cnb_emu_switch(&g_cur->sp, g_sched_sp); }

inline void fiber_exit_bookkeeping() {
    Block* b = g_blk;
    Fiber* f = g_cur;
    f->done = true;
    b->live--;
    if (b->bar_count > 0 && b->bar_count >= b->live) {
        b->bar_count = 0;
        b->bar_gen++;
    }
    Warp& w = b->warps[f->warp];
    w.live--;
    if (w.count > 0 && w.count >= w.live) {
        w.count = 0;
        w.gen++;
    }
}

inline void fiber_entry() {
    (*g_fn)();
    fiber_exit_bookkeeping();
    for (;;) yield();
}

inline void syncthreads() {
    Block* b = g_blk;
    int gen = b->bar_gen;
    if (++b->bar_count >= b->live) {
        b->bar_count = 0;
        b->bar_gen++;
    } else {
        while (b->bar_gen == gen) yield();
    }
}

inline void syncwarp() {
    Warp& w = g_blk->warps[g_cur->warp];
    int gen = w.gen;
    if (++w.count >= w.live) {
        w.count = 0;
        w.gen++;
    } else {
        while (w.gen == gen) yield();
    }
}

template <typename T>
inline T shfl_from(T v, int src_lane) {
    static_assert(sizeof(T) <= 8, "shuffle payload");
    Warp& w = g_blk->warps[g_cur->warp];
    uint64_t bits = 0;
    memcpy(&bits, &v, sizeof(T));
    w.slot[g_cur->lane] = bits;
    syncwarp();
    int src = (src_lane >= 0 && src_lane < 32) ? src_lane : g_cur->lane;
    uint64_t got = w.slot[src];
    syncwarp();
    T r;
    memcpy(&r, &got, sizeof(T));
    return r;
}

inline void run_block(Block& blk, int nthreads, const std::function<void()>& fn, char* stacks, size_t stack_bytes,
                      std::vector<Fiber>& fibers) {
    const int nwarps = (nthreads + 31) / 32;
    blk.live = nthreads;
    blk.bar_count = 0;
    blk.bar_gen = 0;
    blk.warps.assign(nwarps, Warp());
    for (int w = 0; w < nwarps; ++w) blk.warps[w].live = std::min(32, nthreads - w * 32);
    fibers.assign(nthreads, Fiber());
    for (int t = 0; t < nthreads; ++t) {
        Fiber& f = fibers[t];
        f.tidx.x = t % blk.bdim.x;
        f.tidx.y = (t / blk.bdim.x) % blk.bdim.y;
        f.tidx.z = t / (blk.bdim.x * blk.bdim.y);
        f.warp = t / 32;
        f.lane = t % 32;
        uintptr_t top = (uintptr_t)(stacks + (size_t)(t + 1) * stack_bytes);
        top &= ~(uintptr_t)15;
        void** sp = (void**)top;
        *--sp = nullptr;               // alignment pad: rsp % 16 == 8 at fiber_entry
        *--sp = (void*)&fiber_entry;   // return address consumed by `ret`
        for (int i = 0; i < 6; ++i) *--sp = nullptr;  // rbp rbx r12 r13 r14 r15
        f.sp = sp;
    }
    g_blk = &blk;
    g_fn = &fn;
    int ndone = 0;
    while (ndone < nthreads) {
        ndone = 0;
        for (int t = 0; t < nthreads; ++t) {
            if (fibers[t].done) {
                ++ndone;
                continue;
            }
            g_cur = &fibers[t];
            cnb_emu_switch(&g_sched_sp, fibers[t].sp);
            if (fibers[t].done) ++ndone;
        }
    }
    g_cur = nullptr;
}

inline int worker_count() {
    static int n = [] {
        const char* e = getenv("CNB_EMU_THREADS");
        int v = e ? atoi(e) : (int)std::thread::hardware_concurrency();
        return v < 1 ? 1 : (v > 64 ? 64 : v);
    }();
    return n;
}

inline void launch(dim3 grid, dim3 block, size_t smem_bytes, std::function<void()> fn) {
    const long nblocks = (long)grid.x * grid.y * grid.z;
    const int nthreads = (int)(block.x * block.y * block.z);
    if (nblocks <= 0 || nthreads <= 0) return;
    if (nthreads > 1024) {
        fprintf(stderr, "cnb_emu: block of %d threads\n", nthreads);
        abort();
    }
    const size_t stack_bytes = 64 * 1024;
    std::atomic<long> next{0};
    auto worker = [&]() {
        char* stacks = (char*)malloc(stack_bytes * nthreads + 64);
        unsigned char* dsm = smem_bytes ? (unsigned char*)aligned_alloc(128, (smem_bytes + 127) / 128 * 128) : nullptr;
        std::vector<Fiber> fibers;
        Block blk;
        blk.bdim = block;
        blk.gdim = grid;
        blk.dyn_smem = dsm;
        for (;;) {
            long b = next.fetch_add(1);
            if (b >= nblocks) break;
            blk.bidx.x = (unsigned)(b % grid.x);
            blk.bidx.y = (unsigned)((b / grid.x) % grid.y);
            blk.bidx.z = (unsigned)(b / ((long)grid.x * grid.y));
            run_block(blk, nthreads, fn, stacks, stack_bytes, fibers);
        }
        free(stacks);
        free(dsm);
    };
    int nw = (int)std::min<long>(worker_count(), nblocks);
    if (nw <= 1) {
        worker();
    } else {
        std::vector<std::thread> th;
        for (int i = 0; i < nw; ++i) th.emplace_back(worker);
        for (auto& t : th) t.join();
    }
}

}  // namespace cnb_emu

#define threadIdx (cnb_emu::g_cur->tidx)
#define blockIdx (cnb_emu::g_blk->bidx)
#define blockDim (cnb_emu::g_blk->bdim)
#define gridDim (cnb_emu::g_blk->gdim)
#define warpSize 32

inline void __syncthreads() { cnb_emu::syncthreads(); }
inline void __syncwarp(unsigned = 0xffffffffu) { cnb_emu::syncwarp(); }
template <typename T>
inline T __shfl_xor_sync(unsigned, T v, int lane_mask) {
    return cnb_emu::shfl_from(v, cnb_emu::g_cur->lane ^ lane_mask);
}
template <typename T>
inline T __shfl_down_sync(unsigned, T v, int delta) {
    return cnb_emu::shfl_from(v, cnb_emu::g_cur->lane + delta);
}
template <typename T>
inline T __shfl_sync(unsigned, T v, int src) {
    return cnb_emu::shfl_from(v, src);
}

inline float atomicAdd(float* addr, float val) {
    uint32_t* p = reinterpret_cast<uint32_t*>(addr);
    uint32_t old = __atomic_load_n(p, __ATOMIC_RELAXED);
    for (;;) {
        float f;
        memcpy(&f, &old, 4);
        float nf = f + val;
        uint32_t nb;
        memcpy(&nb, &nf, 4);
        if (__atomic_compare_exchange_n(p, &old, nb, false, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) return f;
    }
}
inline double atomicAdd(double* addr, double val) {
    uint64_t* p = reinterpret_cast<uint64_t*>(addr);
    uint64_t old = __atomic_load_n(p, __ATOMIC_RELAXED);
    for (;;) {
        double f;
        memcpy(&f, &old, 8);
        double nf = f + val;
        uint64_t nb;
        memcpy(&nb, &nf, 8);
        if (__atomic_compare_exchange_n(p, &old, nb, false, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) return f;
    }
}
inline int atomicAdd(int* addr, int val) { return __atomic_fetch_add(addr, val, __ATOMIC_RELAXED); }
inline unsigned atomicAdd(unsigned* addr, unsigned val) { return __atomic_fetch_add(addr, val, __ATOMIC_RELAXED); }

template <typename T>
inline T __ldg(const T* p) {
    return *p;
}
inline float rsqrtf(float x) { return 1.0f / sqrtf(x); }
inline float __fdiv_rn(float a, float b) { return a / b; }
inline float __fmul_rn(float a, float b) { return a * b; }
inline float __fsub_rn(float a, float b) { return a - b; }
inline float __fmaf_rn(float a, float b, float c) { return fmaf(a, b, c); }
inline float __frcp_rn(float a) { return 1.0f / a; }
inline unsigned __float_as_uint(float a) {
    unsigned u;
    memcpy(&u, &a, 4);
    return u;
}
inline void __threadfence() { std::atomic_thread_fence(std::memory_order_seq_cst); }

inline void cnb_count_launch();
#define CNB_LAUNCH(kfn, grid, block, smem, stream, ...) \
    (cnb_count_launch(), cnb_emu::launch(dim3(grid), dim3(block), (size_t)(smem), [=]() { kfn(__VA_ARGS__); }))
#define CNB_DYN_SMEM(name) unsigned char* name = cnb_emu::g_blk->dyn_smem
