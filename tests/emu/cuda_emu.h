// cuda_emu.h -- TEST INFRASTRUCTURE. A CUDA-device-language shim that lets g++ compile the product's kernel
// SOURCE (partitionedhashjoin_b200/csrc/phj_kernels.cuh, phj_dist_kernels.cuh, phj_hash.cuh) for the host, so that
// `pytest -m "not gpu"` can execute the kernels' logic on a box without a GPU (tests/test_emulated.py).
//
// It is NOT a CPU fallback of the product: nothing under partitionedhashjoin_b200/ includes, links or loads it,
// the product library still fails with PHJ_ERR_CUDA without a device (tests/test_host.py::test_no_cpu_fallback),
// and no benchmark number comes from here. It models just enough of the execution model to run these kernels:
//
//   * every CUDA thread of a CTA is a fiber on one OS thread (emu_core.cpp); CTAs of a launch run one after the
//     other in blockIdx order (so a CTA may wait only for CTAs with a lower index that have finished -- which is
//     what scan_lookback's ticket order guarantees);
//   * __syncthreads / __syncwarp / votes / shuffles / match are rendezvous points of the live threads of the CTA /
//     warp; a fiber runs until it reaches one. Exited threads count as arrived. All collectives are assumed to name
//     every live lane of the warp (the kernels use converged loops and full masks);
//   * __shared__ is thread_local storage of the executing OS thread; global / "peer" memory is host memory;
//   * atomics are host atomics (the ranks of an emulated multi-GPU join run on several OS threads);
//   * inline PTX lives only in the helper block of phj_kernels.cuh, which has host versions under PHJ_EMULATE.
//
// What it cannot show: timing, memory-model races between threads of a warp, anything about ptxas.
#pragma once
#include <stdint.h>
#include <string.h>

#include <cmath>
#include <type_traits>

#define PHJ_EMULATE 1
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline __attribute__((always_inline))
#define __launch_bounds__(...)
#define __shared__ thread_local  // block scope: implies static; `extern __shared__` names phj::smem_raw
#define __align__(n) __attribute__((aligned(n)))

struct uint2 {
    unsigned int x, y;
};
struct __attribute__((aligned(16))) ulonglong2 {
    unsigned long long x, y;
};
struct uint3 {
    unsigned int x, y, z;
};
inline ulonglong2 make_ulonglong2(unsigned long long x, unsigned long long y) { return ulonglong2{x, y}; }
inline uint2 make_uint2(unsigned int x, unsigned int y) { return uint2{x, y}; }
struct dim3 {
    unsigned int x, y, z;
    constexpr dim3(unsigned int x_ = 1, unsigned int y_ = 1, unsigned int z_ = 1) : x(x_), y(y_), z(z_) {}
};

namespace emu {
struct Ctx {
    uint3 tid, bid;
    dim3 bdim, gdim;
    unsigned lane;
};
extern thread_local Ctx ctx;  // of the fiber that is running on this OS thread

// Rendezvous of the live lanes of the caller's warp: every lane contributes `v`; returns all 32 contributions
// (valid until the caller's next collective) and, in *live, the lanes that took part.
const uint64_t* warp_gather(uint64_t v, uint32_t* live);
void cta_barrier();
// cp.async.bulk shared -> global, modelled as LATE as the PTX rules allow: the copy is only recorded at issue and
// carried out when the issuing thread waits for its bulk groups (wait_group[.read] 0) -- so a kernel that reuses the
// shared-memory source before that wait writes wrong data here, as it may on the device. A thread that exits with
// copies still pending aborts the run (the stores of an exited CTA are not guaranteed to have been read).
void bulk_defer(void* gdst, const void* ssrc, uint32_t bytes);
void bulk_complete();
}  // namespace emu

#define threadIdx (emu::ctx.tid)
#define blockIdx (emu::ctx.bid)
#define blockDim (emu::ctx.bdim)
#define gridDim (emu::ctx.gdim)
constexpr int warpSize = 32;

// ---- barriers, votes, shuffles ---------------------------------------------------------------------
inline void __syncthreads() { emu::cta_barrier(); }
inline void __syncwarp(unsigned = 0xffffffffu) {
    uint32_t live;
    emu::warp_gather(0, &live);
}
inline unsigned __ballot_sync(unsigned mask, int pred) {
    uint32_t live, out = 0;
    const uint64_t* v = emu::warp_gather(pred ? 1 : 0, &live);
    for (int l = 0; l < 32; ++l)
        if ((live >> l & 1) && v[l]) out |= 1u << l;
    return out & mask;
}
inline int __any_sync(unsigned mask, int pred) { return __ballot_sync(mask, pred) != 0; }
inline int __all_sync(unsigned mask, int pred) {
    uint32_t live, out = 0;
    const uint64_t* v = emu::warp_gather(pred ? 1 : 0, &live);
    for (int l = 0; l < 32; ++l)
        if ((live >> l & 1) && v[l]) out |= 1u << l;
    return (out & mask & live) == (mask & live);
}
template <typename T>
inline unsigned __match_any_sync(unsigned mask, T value) {
    static_assert(sizeof(T) <= 8, "");
    uint64_t bits = 0;
    memcpy(&bits, &value, sizeof(T));
    uint32_t live, out = 0;
    const uint64_t* v = emu::warp_gather(bits, &live);
    for (int l = 0; l < 32; ++l)
        if ((live >> l & 1) && v[l] == bits) out |= 1u << l;
    return out & mask;
}
namespace emu {
template <typename T>
inline T pick(const uint64_t* v, uint32_t live, int src, T own) {
    if (src < 0 || src > 31 || !(live >> src & 1)) return own;  // reading an exited lane: undefined on the device
    T out;
    memcpy(&out, &v[src], sizeof(T));
    return out;
}
template <typename T>
inline const uint64_t* share(T value, uint32_t* live) {
    static_assert(sizeof(T) <= 8 && std::is_trivially_copyable<T>::value, "");
    uint64_t bits = 0;
    memcpy(&bits, &value, sizeof(T));
    return warp_gather(bits, live);
}
}  // namespace emu
template <typename T>
inline T __shfl_sync(unsigned, T value, int src, int width = 32) {
    uint32_t live;
    const uint64_t* v = emu::share(value, &live);
    const int lane = (int)emu::ctx.lane, base = lane & ~(width - 1);
    return emu::pick<T>(v, live, base + (src & (width - 1)), value);
}
template <typename T>
inline T __shfl_up_sync(unsigned, T value, unsigned delta, int width = 32) {
    uint32_t live;
    const uint64_t* v = emu::share(value, &live);
    const int lane = (int)emu::ctx.lane, base = lane & ~(width - 1);
    return lane - (int)delta >= base ? emu::pick<T>(v, live, lane - (int)delta, value) : value;
}
template <typename T>
inline T __shfl_down_sync(unsigned, T value, unsigned delta, int width = 32) {
    uint32_t live;
    const uint64_t* v = emu::share(value, &live);
    const int lane = (int)emu::ctx.lane, base = lane & ~(width - 1);
    return lane + (int)delta < base + width ? emu::pick<T>(v, live, lane + (int)delta, value) : value;
}
template <typename T>
inline T __shfl_xor_sync(unsigned, T value, int lane_mask, int width = 32) {
    uint32_t live;
    const uint64_t* v = emu::share(value, &live);
    const int lane = (int)emu::ctx.lane, src = lane ^ lane_mask;
    return (src & ~(width - 1)) == (lane & ~(width - 1)) ? emu::pick<T>(v, live, src, value) : value;
}

// ---- integer intrinsics ----------------------------------------------------------------------------
inline int __popc(unsigned x) { return __builtin_popcount(x); }
inline int __popcll(unsigned long long x) { return __builtin_popcountll(x); }
inline int __ffs(int x) { return __builtin_ffs(x); }
inline int __ffsll(long long x) { return __builtin_ffsll(x); }
inline int __clz(int x) { return x ? __builtin_clz((unsigned)x) : 32; }
inline int __clzll(long long x) { return x ? __builtin_clzll((unsigned long long)x) : 64; }
inline unsigned __brev(unsigned x) {
    unsigned r = 0;
    for (int i = 0; i < 32; ++i) r |= (x >> i & 1u) << (31 - i);
    return r;
}
inline unsigned __umulhi(unsigned a, unsigned b) { return (unsigned)(((unsigned long long)a * b) >> 32); }
inline unsigned long long __umul64hi(unsigned long long a, unsigned long long b) {
    return (unsigned long long)(((unsigned __int128)a * b) >> 64);
}
template <typename T>
inline T __ldg(const T* p) {
    return *p;
}
inline void __threadfence() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }
inline void __threadfence_block() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }

// ---- atomics (return the old value) ----------------------------------------------------------------
template <typename T>
inline T emu_atomic_add(T* p, T v) {
    return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST);
}
inline unsigned atomicAdd(unsigned* p, unsigned v) { return emu_atomic_add(p, v); }
inline int atomicAdd(int* p, int v) { return emu_atomic_add(p, v); }
inline unsigned long long atomicAdd(unsigned long long* p, unsigned long long v) { return emu_atomic_add(p, v); }
template <typename T>
inline T emu_atomic_cas(T* p, T expected, T desired) {
    __atomic_compare_exchange_n(p, &expected, desired, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST);
    return expected;
}
inline unsigned atomicCAS(unsigned* p, unsigned c, unsigned v) { return emu_atomic_cas(p, c, v); }
inline int atomicCAS(int* p, int c, int v) { return emu_atomic_cas(p, c, v); }
inline unsigned long long atomicCAS(unsigned long long* p, unsigned long long c, unsigned long long v) {
    return emu_atomic_cas(p, c, v);
}
inline unsigned atomicExch(unsigned* p, unsigned v) { return __atomic_exchange_n(p, v, __ATOMIC_SEQ_CST); }
inline unsigned long long atomicExch(unsigned long long* p, unsigned long long v) {
    return __atomic_exchange_n(p, v, __ATOMIC_SEQ_CST);
}
inline unsigned atomicOr(unsigned* p, unsigned v) { return __atomic_fetch_or(p, v, __ATOMIC_SEQ_CST); }
inline unsigned atomicMax(unsigned* p, unsigned v) {
    unsigned old = *p;
    while (old < v && !__atomic_compare_exchange_n(p, &old, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST)) {
    }
    return old;
}
inline unsigned long long atomicMax(unsigned long long* p, unsigned long long v) {
    unsigned long long old = *p;
    while (old < v && !__atomic_compare_exchange_n(p, &old, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST)) {
    }
    return old;
}
inline unsigned atomicInc(unsigned* p, unsigned limit) {  // old >= limit ? 0 : old + 1
    unsigned old = *p;
    while (!__atomic_compare_exchange_n(p, &old, old >= limit ? 0u : old + 1u, false, __ATOMIC_SEQ_CST,
                                        __ATOMIC_SEQ_CST)) {
    }
    return old;
}

// ---- min / max as CUDA overloads them ---------------------------------------------------------------
#define PHJ_EMU_MINMAX(T)                        \
    inline T min(T a, T b) { return b < a ? b : a; } \
    inline T max(T a, T b) { return a < b ? b : a; }
PHJ_EMU_MINMAX(int)
PHJ_EMU_MINMAX(unsigned)
PHJ_EMU_MINMAX(long)
PHJ_EMU_MINMAX(unsigned long)
PHJ_EMU_MINMAX(long long)
PHJ_EMU_MINMAX(unsigned long long)
PHJ_EMU_MINMAX(double)
#undef PHJ_EMU_MINMAX
inline unsigned min(unsigned a, int b) { return min(a, (unsigned)b); }
inline unsigned min(int a, unsigned b) { return min((unsigned)a, b); }
inline unsigned long min(unsigned long a, unsigned b) { return min(a, (unsigned long)b); }
inline unsigned long min(unsigned a, unsigned long b) { return min((unsigned long)a, b); }
inline unsigned long long min(unsigned long long a, unsigned long b) { return min(a, (unsigned long long)b); }
inline unsigned long long min(unsigned long a, unsigned long long b) { return min((unsigned long long)a, b); }
inline unsigned long max(unsigned long a, unsigned b) { return max(a, (unsigned long)b); }
inline unsigned long max(unsigned a, unsigned long b) { return max((unsigned long)a, b); }

using std::ceil;
using std::exp;
using std::floor;
using std::log;
using std::pow;
