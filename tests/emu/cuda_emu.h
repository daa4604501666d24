// cuda_emu.h -- TEST INFRASTRUCTURE: a single-OS-thread functional emulator of the CUDA subset the
// CUDA-core kernels of this repository use (1-D thread blocks, __syncthreads and its count/or forms,
// full-mask warp collectives, shared-memory atomics, streaming loads, IEEE intrinsics).
//
// Every CUDA thread of a block is a ucontext fiber; a barrier or a warp collective yields to the next
// fiber until the rendezvous completes.  Fibers run in thread order and only switch at barriers /
// collectives / spin waits, which is an adversarial schedule for a missing __syncthreads (thread 0 runs a
// whole phase before thread 1 starts), and makes every run deterministic.  A rendezvous that cannot
// complete (divergent barrier) is reported as a deadlock instead of hanging.
//
// Nothing under vietnamese_qa_system_b200/ includes this file; tests/emu/build.py compiles the product's
// kernel headers against it (after a mechanical source transform of `extern __shared__` declarations and
// the three inline-PTX helpers) into tests/emu/_build/libvqa_emu.so for the CPU tests.
#pragma once
#include <cuda_runtime.h>  // vector types, dim3, qualifiers as host no-ops
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <ucontext.h>

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <stdexcept>
#include <string>
#include <vector>

#undef __shared__
#define __shared__ static
#undef __launch_bounds__
#define __launch_bounds__(...)
#ifndef __noinline__
#define __noinline__ __attribute__((noinline))
#endif

namespace emu {

constexpr int kMaxThreads = 1024;
constexpr size_t kStackBytes = 256 * 1024;

struct Fiber {
    ucontext_t ctx;
    char *stack = nullptr;
    bool done = true;
};

constexpr int kMaxCluster = 4;

struct CtaBarrier {  // __syncthreads and its count / or forms, per CTA
    int arrived = 0, gen = 0, acc_cnt = 0, acc_or = 0, res_cnt = 0, res_or = 0, alive = 0;
};

struct State {
    Fiber f[kMaxThreads];
    ucontext_t sched;
    int cur = 0, n = 0;            // running fiber, fibers in flight (= cluster * nthr)
    int nthr = 0, cluster = 1;     // threads per CTA, CTAs that run together (a thread-block cluster)
    std::function<void()> body;
    CtaBarrier bar[kMaxCluster];
    int cl_arrived = 0, cl_gen = 0;  // barrier.cluster
    // warp collectives (warps never straddle CTAs: block sizes are multiples of 32)
    int w_arrived[kMaxThreads / 32] = {}, w_gen[kMaxThreads / 32] = {};
    unsigned long long wbuf[kMaxThreads / 32][32] = {};
    long progress = 0;
    std::string error;
    // fiber order inside a scheduling pass: 0 thread order, 1 reverse, 2 a fresh random permutation per pass
    int schedule = 0;
    unsigned long long rng = 0x9E3779B97F4A7C15ull;
    int order[kMaxThreads];
    // dynamic shared memory of every CTA of the running cluster
    // (1024-aligned: offsets and addresses share their low bits, which the 128-byte swizzle works on)
    alignas(1024) unsigned char dyn_smem[kMaxCluster][232 * 1024];
};
inline State &S() {
    static State s;
    return s;
}
inline int cur_cta() { return S().nthr ? S().cur / S().nthr : 0; }
inline unsigned char *cur_smem() { return S().dyn_smem[cur_cta()]; }

inline void yield() {
    State &s = S();
    swapcontext(&s.f[s.cur].ctx, &s.sched);
}

inline void complete_barrier(State &s, CtaBarrier &b) {
    b.res_cnt = b.acc_cnt;
    b.res_or = b.acc_or;
    b.acc_cnt = b.acc_or = 0;
    b.arrived = 0;
    ++b.gen;
    ++s.progress;
}

inline void block_barrier(int pred) {
    State &s = S();
    CtaBarrier &b = s.bar[cur_cta()];
    const int gen = b.gen;
    b.acc_cnt += pred != 0;
    b.acc_or |= pred != 0;
    if (++b.arrived == b.alive) complete_barrier(s, b);
    else
        while (b.gen == gen) yield();
}

inline void cluster_barrier() {  // all threads of all CTAs of the cluster
    State &s = S();
    const int gen = s.cl_gen;
    if (++s.cl_arrived == s.n) {
        s.cl_arrived = 0;
        ++s.cl_gen;
        ++s.progress;
    } else
        while (s.cl_gen == gen) yield();
}

inline void warp_barrier() {
    State &s = S();
    const int w = s.cur >> 5;
    const int gen = s.w_gen[w];
    if (++s.w_arrived[w] == 32) {
        s.w_arrived[w] = 0;
        ++s.w_gen[w];
        ++s.progress;
    } else
        while (s.w_gen[w] == gen) yield();
}

template <typename T>
inline void warp_gather(T v, T (&all)[32]) {
    static_assert(sizeof(T) <= 8, "warp collectives carry at most 8 bytes");
    State &s = S();
    const int w = s.cur >> 5, l = s.cur & 31;
    s.wbuf[w][l] = 0;
    std::memcpy(&s.wbuf[w][l], &v, sizeof(T));
    warp_barrier();
    for (int i = 0; i < 32; ++i) std::memcpy(&all[i], &s.wbuf[w][i], sizeof(T));
    warp_barrier();
}

inline void trampoline() {
    State &s = S();
    try {
        s.body();
    } catch (const std::exception &e) {  // an exception must not unwind past the fiber's first frame
        if (s.error.empty()) s.error = e.what();
    }
    s.f[s.cur].done = true;
}

}  // namespace emu

// ---- built-in variables (1-D blocks; 2-D grids) ----------------------------------------------------
inline uint3 threadIdx, blockIdx;
inline dim3 blockDim, gridDim;

namespace emu {

// Run one cluster of `cluster` thread blocks (blockIdx.x = first_bx .. first_bx + cluster - 1) together: `body` is
// executed once per CUDA thread; a plain launch is a cluster of one.
inline void run_cluster(unsigned nthreads, unsigned first_bx, unsigned by, int cluster, const std::function<void()> &body) {
    State &s = S();
    if (nthreads == 0 || nthreads % 32 != 0 || cluster < 1 || cluster > kMaxCluster ||
        nthreads * (unsigned)cluster > (unsigned)kMaxThreads)
        throw std::runtime_error("emu: block size must be a multiple of 32, cluster <= 4, <= 1024 threads in flight");
    s.nthr = (int)nthreads;
    s.cluster = cluster;
    s.n = (int)nthreads * cluster;
    s.body = body;
    for (int c = 0; c < cluster; ++c) {
        s.bar[c] = CtaBarrier();
        s.bar[c].alive = (int)nthreads;
    }
    s.cl_arrived = 0;
    for (auto &x : s.w_arrived) x = 0;
    blockIdx.y = by;
    blockIdx.z = 0;
    blockDim = dim3(nthreads, 1, 1);
    for (int t = 0; t < s.n; ++t) {
        Fiber &f = s.f[t];
        if (!f.stack) f.stack = static_cast<char *>(std::malloc(kStackBytes));
        getcontext(&f.ctx);
        f.ctx.uc_stack.ss_sp = f.stack;
        f.ctx.uc_stack.ss_size = kStackBytes;
        f.ctx.uc_link = &s.sched;
        makecontext(&f.ctx, trampoline, 0);
        f.done = false;
    }
    int remaining = s.n;
    int idle_rounds = 0;
    for (int t = 0; t < s.n; ++t) s.order[t] = s.schedule == 1 ? s.n - 1 - t : t;
    while (remaining > 0) {
        const long before = s.progress;
        if (s.schedule == 2) {  // Fisher-Yates with xorshift64*: a different interleaving on every pass
            for (int t = s.n - 1; t > 0; --t) {
                s.rng ^= s.rng >> 12;
                s.rng ^= s.rng << 25;
                s.rng ^= s.rng >> 27;
                const int j = (int)((s.rng * 0x2545F4914F6CDD1Dull >> 33) % (unsigned)(t + 1));
                const int tmp = s.order[t];
                s.order[t] = s.order[j];
                s.order[j] = tmp;
            }
        }
        for (int oi = 0; oi < s.n; ++oi) {
            const int t = s.order[oi];
            if (s.f[t].done) continue;
            s.cur = t;
            threadIdx.x = (unsigned)(t % s.nthr);
            threadIdx.y = threadIdx.z = 0;
            blockIdx.x = first_bx + (unsigned)(t / s.nthr);
            swapcontext(&s.sched, &s.f[t].ctx);
            if (!s.error.empty()) {  // abandon the cluster: the other fibers' stacks are simply dropped
                const std::string msg = s.error;
                s.error.clear();
                for (int u = 0; u < s.n; ++u) s.f[u].done = true;
                throw std::runtime_error(msg);
            }
            if (s.f[t].done) {
                --remaining;
                ++s.progress;
                // exited threads no longer take part in their CTA's barriers
                CtaBarrier &b = s.bar[t / s.nthr];
                --b.alive;
                if (b.alive > 0 && b.arrived == b.alive) complete_barrier(s, b);
            }
        }
        if (s.progress == before) {
            if (++idle_rounds > 4) {
                for (int u = 0; u < s.n; ++u) s.f[u].done = true;
                throw std::runtime_error("emu: deadlock (divergent barrier, warp collective or mbarrier wait)");
            }
        } else
            idle_rounds = 0;
    }
}

constexpr size_t kMaxDynSmem = 227 * 1024;  // what a CTA can opt in to on sm_100

template <typename F>
inline void launch(dim3 grid, unsigned nthreads, size_t dyn_smem_bytes, const F &body, int cluster = 1) {
    if (dyn_smem_bytes > kMaxDynSmem)
        throw std::runtime_error("emu: launch asks for " + std::to_string(dyn_smem_bytes) +
                                 " bytes of dynamic shared memory (limit " + std::to_string(kMaxDynSmem) + ")");
    if (cluster < 1 || cluster > kMaxCluster || grid.x % (unsigned)cluster != 0)
        throw std::runtime_error("emu: grid.x must be a multiple of the cluster size (<= 4)");
    gridDim = grid;
    State &s = S();
    // everything past the bytes the launch asked for is a guard zone: a kernel whose shared-memory layout
    // outgrows its size formula is caught here instead of silently corrupting (or faulting) on the device
    const size_t guard_len = sizeof(s.dyn_smem[0]) - dyn_smem_bytes;
    for (unsigned by = 0; by < grid.y; ++by)
        for (unsigned bx = 0; bx < grid.x; bx += (unsigned)cluster) {
            for (int c = 0; c < cluster; ++c) std::memset(s.dyn_smem[c] + dyn_smem_bytes, 0xCB, guard_len);
            run_cluster(nthreads, bx, by, cluster, body);
            for (int c = 0; c < cluster; ++c)
                for (size_t i = 0; i < guard_len; ++i)
                    if (s.dyn_smem[c][dyn_smem_bytes + i] != 0xCB)
                        throw std::runtime_error("emu: shared memory written at byte " +
                                                 std::to_string(dyn_smem_bytes + i) + " but the launch asked for only " +
                                                 std::to_string(dyn_smem_bytes));
        }
}

}  // namespace emu

// ---- synchronisation --------------------------------------------------------------------------------
inline void __syncthreads() { emu::block_barrier(0); }
inline int __syncthreads_count(int pred) {
    emu::block_barrier(pred);
    return emu::S().bar[emu::cur_cta()].res_cnt;
}
inline int __syncthreads_or(int pred) {
    emu::block_barrier(pred);
    return emu::S().bar[emu::cur_cta()].res_or;
}
inline void __syncwarp(unsigned = 0xffffffffu) { emu::warp_barrier(); }
inline void __threadfence_block() {}
inline void __threadfence() {}
inline void __threadfence_system() {}
[[noreturn]] inline void __trap() { throw std::runtime_error("emu: __trap()"); }
inline void __nanosleep(unsigned) { emu::yield(); }  // spin waits make progress by letting the others run

// ---- warp collectives (full mask only) --------------------------------------------------------------
inline void emu_check_mask(unsigned m) {
    if (m != 0xffffffffu) throw std::runtime_error("emu: only full-mask warp collectives are emulated");
}
template <typename T>
inline T __shfl_sync(unsigned m, T v, int src, int = 32) {
    emu_check_mask(m);
    T all[32];
    emu::warp_gather(v, all);
    return all[src & 31];
}
template <typename T>
inline T __shfl_xor_sync(unsigned m, T v, int lane_mask, int = 32) {
    emu_check_mask(m);
    T all[32];
    emu::warp_gather(v, all);
    return all[(threadIdx.x & 31) ^ (lane_mask & 31)];
}
template <typename T>
inline T __shfl_up_sync(unsigned m, T v, unsigned delta, int = 32) {
    emu_check_mask(m);
    T all[32];
    emu::warp_gather(v, all);
    const int l = threadIdx.x & 31;
    return l >= (int)delta ? all[l - delta] : v;
}
template <typename T>
inline T __shfl_down_sync(unsigned m, T v, unsigned delta, int = 32) {
    emu_check_mask(m);
    T all[32];
    emu::warp_gather(v, all);
    const int l = threadIdx.x & 31;
    return l + (int)delta < 32 ? all[l + delta] : v;
}
inline unsigned __ballot_sync(unsigned m, int pred) {
    emu_check_mask(m);
    int all[32];
    emu::warp_gather(pred != 0 ? 1 : 0, all);
    unsigned r = 0;
    for (int i = 0; i < 32; ++i) r |= (unsigned)all[i] << i;
    return r;
}
// event counter for tests (tests/emu/build.py injects increments into selected device functions)
inline long long &emu_event_counter() {
    static long long c = 0;
    return c;
}
inline unsigned __match_any_sync(unsigned m, unsigned v) {
    emu_check_mask(m);
    unsigned all[32];
    emu::warp_gather(v, all);
    unsigned r = 0;
    for (int i = 0; i < 32; ++i) r |= (unsigned)(all[i] == v) << i;
    return r;
}
inline int __reduce_max_sync(unsigned m, int v) {
    emu_check_mask(m);
    int all[32];
    emu::warp_gather(v, all);
    int r = all[0];
    for (int i = 1; i < 32; ++i) r = all[i] > r ? all[i] : r;
    return r;
}
inline int __popc(unsigned v) { return __builtin_popcount(v); }
inline int __ffs(int v) { return __builtin_ffs(v); }
inline int __clz(int v) { return v == 0 ? 32 : __builtin_clz((unsigned)v); }

// ---- atomics (cooperative fibers: plain read-modify-write) --------------------------------------------
template <typename T>
inline T atomicAdd(T *p, T v) {
    T o = *p;
    *p = o + v;
    return o;
}
template <typename T>
inline T atomicMax(T *p, T v) {
    T o = *p;
    if (v > o) *p = v;
    return o;
}
template <typename T>
inline T atomicOr(T *p, T v) {
    T o = *p;
    *p = o | v;
    return o;
}
template <typename T>
inline T atomicCAS(T *p, T cmp, T v) {
    T o = *p;
    if (o == cmp) *p = v;
    return o;
}
template <typename T>
inline T atomicExch(T *p, T v) {
    T o = *p;
    *p = v;
    return o;
}

// ---- loads and IEEE intrinsics (compile with -ffp-contract=off) ----------------------------------------
template <typename T>
inline T __ldg(const T *p) { return *p; }
inline float __fadd_rn(float a, float b) { return a + b; }
inline float __fmul_rn(float a, float b) { return a * b; }
inline float __fmaf_rn(float a, float b, float c) { return std::fmaf(a, b, c); }  // single rounding, like FFMA
inline double __dadd_rn(double a, double b) { return a + b; }
inline double __dmul_rn(double a, double b) { return a * b; }
inline double __ddiv_rn(double a, double b) { return a / b; }
inline float __double2float_rn(double a) { return (float)a; }
inline unsigned __float_as_uint(float f) { unsigned u; std::memcpy(&u, &f, 4); return u; }
inline float __uint_as_float(unsigned u) { float f; std::memcpy(&f, &u, 4); return f; }
inline int __float_as_int(float f) { int u; std::memcpy(&u, &f, 4); return u; }
inline float __int_as_float(int u) { float f; std::memcpy(&f, &u, 4); return f; }
inline double __longlong_as_double(long long v) { double d; std::memcpy(&d, &v, 8); return d; }

inline int min(int a, int b) { return a < b ? a : b; }
inline int max(int a, int b) { return a > b ? a : b; }
inline long long min(long long a, long long b) { return a < b ? a : b; }
inline long long max(long long a, long long b) { return a > b ? a : b; }
inline unsigned min(unsigned a, unsigned b) { return a < b ? a : b; }
inline unsigned max(unsigned a, unsigned b) { return a > b ? a : b; }
