// emu_ptx_helpers.h -- TEST INFRASTRUCTURE: host versions of the inline-PTX helper block of phj_kernels.cuh (the
// block between `#ifndef PHJ_PTX_HELPERS_PROVIDED` and its `#endif`), same names and signatures. Loads and stores
// become plain accesses, the TMA bulk store a copy that happens when its thread waits for its bulk groups (the latest
// moment the PTX rules allow), the proxy fence a no-op, bar.warp.sync a warp rendezvous.
#pragma once
#include <time.h>

#include "cuda_emu.h"

#define PHJ_PTX_HELPERS_PROVIDED 1
#define PHJ_SCAT_PEERS4 0  // warp_peers: the C++ form (the PTX spelling of the same ballots cannot run here)

namespace phj {

inline ulonglong2 ld_stream_v2(const ulonglong2* p) { return *p; }
struct __attribute__((aligned(32))) Bucket4 {
    unsigned long long k0, k1, k2, k3;
};
inline Bucket4 ld_bucket4(const uint64_t* p) {
    Bucket4 b;
    b.k0 = __atomic_load_n(p + 0, __ATOMIC_RELAXED);
    b.k1 = __atomic_load_n(p + 1, __ATOMIC_RELAXED);
    b.k2 = __atomic_load_n(p + 2, __ATOMIC_RELAXED);
    b.k3 = __atomic_load_n(p + 3, __ATOMIC_RELAXED);
    return b;
}
inline uint64_t ld_stream_u64(const uint64_t* p) { return *p; }
inline void st_stream_v2(ulonglong2* p, const ulonglong2& v) { *p = v; }
inline uint32_t lanemask_lt() { return (1u << emu::ctx.lane) - 1u; }
inline uint64_t globaltimer_ns() {
    timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (uint64_t)ts.tv_sec * 1000000000ull + (uint64_t)ts.tv_nsec;
}
inline void cta_sync() {
    __syncwarp();
    __syncthreads();
}
inline void bulk_store_s2g(void* gdst, const void* ssrc, uint32_t bytes) { emu::bulk_defer(gdst, ssrc, bytes); }
inline void bulk_commit() {}
inline void bulk_wait_read0() { emu::bulk_complete(); }
inline void bulk_wait_all0() { emu::bulk_complete(); }
inline void fence_proxy_async_smem() {}

}  // namespace phj
