// phj_engine.cu -- device arenas, launch plan and the C ABI of include/phj.h.
//
// Replaces, for the B200, the orchestration the reference does in
//   RadixClustering::HashJoiner::Run / Partition / Join   src/RadixCluster/HashJoin.hpp:190-440
//   NoPartitioning::HashJoiner::Run / Build / Probe       src/NoPartitioning/HashJoin.hpp:54-187
// The reference's thread pool + 3-stage Pipeline (src/Common/ThreadPool.cpp:99-146) becomes one
// CUDA stream: kernel order on the stream is the pipeline barrier, the grid is the worker set.
// Nothing in the timed region waits for the host: segment plans, cursors and partition boundaries
// are produced and consumed on the device; the host reads back 16 bytes at the end.
#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "../../include/phj.h"

// Tile shapes of the partitioning kernels (tuples per tile = TPB * IPT).
#ifndef PHJ_HIST_TPB
#define PHJ_HIST_TPB 512
#endif
#ifndef PHJ_HIST_IPT
#define PHJ_HIST_IPT 8
#endif
#ifndef PHJ_SCAT_TPB
#define PHJ_SCAT_TPB 512
#endif
#ifndef PHJ_SCAT_IPT
#define PHJ_SCAT_IPT 8
#endif
#ifndef PHJ_SCAT_MINB
#define PHJ_SCAT_MINB 2  // two 512-thread CTAs per SM: caps the scatter at 64 registers (no spills)
#endif
#ifndef PHJ_SCAT_PSCAN
#define PHJ_SCAT_PSCAN 1  // all-warp counter scan in radix_scatter (0: the two-warp scan of round 1, for the A/B)
#endif
#ifndef PHJ_SEGS_PER_SM
#define PHJ_SEGS_PER_SM 16  // segments (= CTAs of the partitioning kernels) per SM for the large relation: 8 -> 16
                            // shortens every kernel's tail (histogram 524 -> 485 us, join 3.92 -> 3.84 ms)
#endif
#ifndef PHJ_SCAT_ALLWRITE
#define PHJ_SCAT_ALLWRITE 1  // all lanes of a digit group store the warp counter (0: an elected leader; measured
                             // 1299 / 1380 -> 1284 / 1362 us per pass, tools/ab_scatter.py)
#endif
#ifndef PHJ_SCAT_MINB_PLAIN
#define PHJ_SCAT_MINB_PLAIN PHJ_SCAT_MINB  // CTAs per SM of the scatter without the fused pass-2 histogram
#endif
#ifndef PHJ_JOIN_WAVES
#define PHJ_JOIN_WAVES 4  // probe slices (CTAs) of join_partitions per resident CTA slot
#endif
#ifndef PHJ_JOIN_TPB
#define PHJ_JOIN_TPB 512
#endif
#ifndef PHJ_JOIN_U
#define PHJ_JOIN_U 4  // probe tuples in flight per thread of join_partitions (tools/tune_shapes.py: j*u6 / u8)
#endif
#ifndef PHJ_JOIN_BUCKET
#define PHJ_JOIN_BUCKET 2  // keys per shared-memory bucket (2: one LDS.128 per probe step)
#endif
#include "phj_kernels.cuh"

namespace {

thread_local std::string g_error;

int fail(int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_error = buf;
    return code;
}

#define PHJ_CUDA(call)                                                                          \
    do {                                                                                        \
        cudaError_t e_ = (call);                                                                \
        if (e_ != cudaSuccess)                                                                  \
            return fail(e_ == cudaErrorMemoryAllocation ? PHJ_ERR_NOMEM : PHJ_ERR_CUDA,         \
                        "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__,       \
                        __LINE__);                                                              \
    } while (0)

constexpr int kScatTile = PHJ_SCAT_TPB * PHJ_SCAT_IPT;
constexpr int kMaxBitsPerPass = 8;
constexpr int kMaxKernelTimes = 48;
constexpr int kMaxSplitDigits = 256;  // PHJ_ALGO_SHARD_SPLIT: owner ranks x local pass-1 digits
constexpr int kMaxSplitChunks = 16;
constexpr int kMaxUploadChunks = 32;                     // streamed phj_join_host
constexpr size_t kUploadChunkBytes = (size_t)256 << 20;  // ... automatic chunking: ~256 MB of probe tuples

enum Scalar : int {  // device-resident uint32 scalars
    kNsegs1 = 0,
    kNcounts1,
    kNsegs2,
    kNcounts2,
    kUnused4,  // (the oversize count and plan_pass2's truncation flag now share a word behind the match counter)
    kUnused5,
    kGtFlags,
    kCtCursor,
    kNumScalars = 8
};

struct KernelTime {
    const char* name;
    cudaEvent_t begin, end;
    bool used;
};

}  // namespace

// One process driving several GPUs (phj_config.num_gpus > 1): defined in phj_dist.inl.
struct phj_group;
int group_create(const phj_config* cfg, phj_group** out);
void group_destroy(phj_group* g);
int group_set_relations(phj_group* g, const void* build, size_t n_build, const void* probe, size_t n_probe,
                        bool device_resident);
int group_join(phj_group* g, phj_result* out);

struct phj_handle {
    phj_config cfg{};
    phj_group* group = nullptr;  // num_gpus > 1: every call on this handle is forwarded to the group
    int device = 0;
    int sm_count = 0;
    size_t smem_optin = 0;
    cudaStream_t stream = nullptr;

    // relations: 0 = build (R), 1 = probe (S)
    ulonglong2* d_in[2] = {nullptr, nullptr};
    bool owns_in[2] = {false, false};
    size_t cap_in[2] = {0, 0};
    size_t n[2] = {0, 0};
    bool have_data = false;
    ulonglong2* d_buf_a[2] = {nullptr, nullptr};  // pass-1 output
    ulonglong2* d_buf_b[2] = {nullptr, nullptr};  // pass-2 output
    size_t cap_buf[2] = {0, 0};

    // radix plan
    bool pow2 = true;
    uint64_t P = 0;     // reference fan-out (hash % P)
    int bits_total = 0; // bits of the full digit space (ceil(log2 P))
    int b1 = 0, b2 = 0;
    uint32_t d1 = 0, d2 = 0;  // digits actually used per pass
    uint64_t nparts = 0;      // d1 * d2 (>= P)
    uint32_t nsegs1 = 0, max_segs2 = 0, target_segs2[2] = {0, 0};
    uint32_t nseg1_rel[2] = {0, 0}, cnt_base1_rel[2] = {0, 0};
    uint64_t seg_len[2] = {0, 0};   // tuples per pass-1 segment
    uint64_t seg_len2[2] = {0, 0};  // tuples per pass-2 segment (shorter: parents end ragged, more and
                                    // shorter segments balance the CTAs better)
    phj::Parent2* d_parents2[2] = {nullptr, nullptr};
    size_t cap_parents2 = 0;
    bool fuse2 = false;             // pass-2 histogram accumulated by the pass-1 scatter
    bool hist12 = false;            // ... or both passes' histograms from one read (radix_histogram_full)
    uint32_t* d_hist12 = nullptr;   // [pass-1 segment][64 x 64]
    size_t cap_hist12 = 0;
    ulonglong2** d_outd[2] = {nullptr, nullptr};  // shard split: per-owner destination bases
    bool shard_counted = false;
    uint32_t nchunks = 1;                         // row chunks of the probe relation
    uint32_t chunk_first_seg[2][kMaxSplitChunks + 1] = {};  // per relation, in its own segment numbering
    uint64_t* d_shard_starts = nullptr;           // [rel][digit][chunk + 1] (split_starts kernel)
    uint64_t* h_shard_starts = nullptr;           // pinned copy
    size_t plan_n[2] = {0, 0};                    // relation sizes the current plan has room for
    size_t segs_n[2] = {0, 0};                    // relation sizes the pass-1 segment table covers
    // relations bound already partitioned by the pass-1 digit (phj_bind_device_partitioned)
    bool prepart = false;
    uint32_t prepart_parents = 0;      // plan built for this many pre-partitioned parents (0 = the plan's own)
    uint32_t prepart_first = 0;        // full-digit-space index of the first parent bound
    uint64_t prepart_space = 0;        // size of that digit space (a power of two; 0 = this plan's 2^b1)
    std::vector<uint32_t> parent_digits;  // optional: digit of every bound parent (else first + index)
    uint64_t* d_pre_bounds = nullptr;  // [2][d1 + 1]
    size_t cap_pre_bounds = 0;
    phj::Segment* d_segs1 = nullptr;
    phj::Segment* d_segs2 = nullptr;
    size_t cap_segs1 = 0, cap_segs2 = 0;
    uint32_t* d_scalars = nullptr;
    uint32_t* d_counts = nullptr;
    uint64_t* d_cursors = nullptr;
    phj::ScanState* d_scan_state = nullptr;  // scan_lookback: chunk sums + epochs + ticket
    uint32_t scan_epoch = 0;
    size_t cap_counts = 0, cap_cursors = 0;
    uint64_t* d_bounds1[2] = {nullptr, nullptr};
    uint64_t* d_bounds2[2] = {nullptr, nullptr};
    size_t cap_bounds1 = 0, cap_bounds2 = 0;

    // join
    unsigned long long* d_matches = nullptr;
    uint64_t* d_cta_times = nullptr;
    size_t cap_cta_times = 0;
    uint32_t join_grid = 0, join_slots = 0, join_max_keys = 0;
    uint64_t* d_gt = nullptr;
    uint64_t gt_buckets = 0;
    // PHJ_FLAG_L2_TABLES: per-partition tables in global memory, probed out of L2 (pt_build / pt_probe)
    bool l2join = false;
    uint64_t* d_pt = nullptr;
    size_t cap_pt = 0;
    uint32_t pt_region_buckets = 0;
    // bucket-chained table (PHJ_FLAG_CHAINED_TABLE)
    uint32_t* d_ct_heads = nullptr;
    phj::ChainBucket* d_ct_buckets = nullptr;
    uint32_t ct_nheads = 0, ct_pool = 0;

    // joined table (phj_join_materialize)
    int64_t* d_joined = nullptr;
    uint64_t cap_joined = 0, n_joined = 0;
    unsigned long long* d_cta_rows = nullptr;  // rows per CTA of join_materialize, then their scan

    // host staging (pinned)
    uint64_t* h_out = nullptr;  // [0] matches, [1] scalars copy...
    uint64_t* h_cta_times = nullptr;

    // streamed phj_join_host: the probe relation is uploaded in row chunks on `upload_stream` while
    // `stream_child` joins the chunks that have landed against the whole build relation
    phj_handle* stream_child = nullptr;
    cudaStream_t upload_stream = nullptr;
    cudaEvent_t upload_ev[kMaxUploadChunks + 2] = {};  // [0] begin, [1] build landed, [2 + c] chunk c landed

    cudaEvent_t ev[6] = {};
    KernelTime ktimes[kMaxKernelTimes] = {};
    int n_ktimes = 0;
    bool time_kernels = false;
    std::string ktime_filter;  // non-empty: only kernels whose scope name contains it are timed
    bool use_match = false;   // PHJ_RANK=match: match.any instead of ballots in the stable kernels
    bool use_lanes = true;
    bool lanes_scatter = false;    // PHJ_RANK=stable|match disables the lane-private kernels
    bool debug_sync = false;  // PHJ_DEBUG_SYNC=1: synchronise and check after every launch
    bool joined_radix = false;
    uint32_t launches = 0;
};

namespace {

using namespace phj;

int ilog2_ceil(uint64_t x) {
    int b = 0;
    while ((1ull << b) < x) ++b;
    return b;
}

template <typename T>
int dev_reserve(T** p, size_t* cap, size_t want) {
    if (want <= *cap && *p) return PHJ_OK;
    if (*p) cudaFree(*p);
    *p = nullptr;
    *cap = 0;
    if (want == 0) want = 1;
    PHJ_CUDA(cudaMalloc(reinterpret_cast<void**>(p), want * sizeof(T)));
    *cap = want;
    return PHJ_OK;
}

// ---- kernel-time bookkeeping --------------------------------------------------------------------
struct KernelScope {
    phj_handle* h;
    int idx;
    const char* name_;
    cudaStream_t stream;
    KernelScope(phj_handle* h_, const char* name, uint32_t kernels = 1, cudaStream_t on = nullptr)
        : h(h_), idx(-1), name_(name), stream(on ? on : h_->stream) {
        h->launches += kernels;  // kernels launched inside this scope (phj_result.kernel_launches)
        if (h->time_kernels && h->n_ktimes < kMaxKernelTimes &&
            (h->ktime_filter.empty() || strstr(name, h->ktime_filter.c_str()))) {
            idx = h->n_ktimes++;
            h->ktimes[idx].name = name;
            h->ktimes[idx].used = true;
            cudaEventRecord(h->ktimes[idx].begin, stream);
        }
    }
    ~KernelScope() {
        if (idx >= 0) cudaEventRecord(h->ktimes[idx].end, stream);
        if (h->debug_sync) {
            cudaError_t e = cudaStreamSynchronize(h->stream);
            if (e == cudaSuccess) e = cudaGetLastError();
            if (e != cudaSuccess)
                fprintf(stderr, "[phj debug] kernel %s (launch #%u) failed: %s\n", name_, h->launches,
                        cudaGetErrorString(e));
        }
    }
};

// ---- template dispatch ---------------------------------------------------------------------------
template <int BITS, int HASH, bool POW2>
void launch_hist_t(phj_handle* h, const PassParams& pp, uint32_t grid) {
    if (h->use_match)
        radix_histogram<BITS, HASH, POW2, PHJ_HIST_TPB, PHJ_HIST_IPT, false>
            <<<grid, PHJ_HIST_TPB, 0, h->stream>>>(pp);
    else
        radix_histogram<BITS, HASH, POW2, PHJ_HIST_TPB, PHJ_HIST_IPT, true>
            <<<grid, PHJ_HIST_TPB, 0, h->stream>>>(pp);
}

template <int BITS, int HASH, bool POW2, bool TMA, bool BALLOT, bool FUSE2>
cudaError_t launch_scatter_tbf(phj_handle* h, const PassParams& pp, uint32_t grid) {
    using L = ScatterSmem<BITS, PHJ_SCAT_TPB, PHJ_SCAT_IPT>;
    size_t smem = L::bytes(TMA) + (FUSE2 ? L::fuse2_bytes : 0);
    auto kern = radix_scatter<BITS, HASH, POW2, PHJ_SCAT_TPB, PHJ_SCAT_IPT, TMA, BALLOT, FUSE2,
                              (FUSE2 || !TMA) ? PHJ_SCAT_MINB : PHJ_SCAT_MINB_PLAIN>;
    static bool configured[16] = {};
    if (!configured[h->device & 15]) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        configured[h->device & 15] = true;
    }
    kern<<<grid, PHJ_SCAT_TPB, smem, h->stream>>>(pp);
    return cudaSuccess;
}

template <int BITS, int HASH, bool POW2, bool TMA, bool BALLOT>
cudaError_t launch_scatter_tb(phj_handle* h, const PassParams& pp, uint32_t grid) {
    if (BITS == 6 && pp.counts2 != nullptr)  // fused pass-2 histogram (6-bit passes only)
        return launch_scatter_tbf<6, HASH, POW2, TMA, BALLOT, true>(h, pp, grid);
    return launch_scatter_tbf<BITS, HASH, POW2, TMA, BALLOT, false>(h, pp, grid);
}

template <int BITS, int HASH, bool POW2, bool TMA>
cudaError_t launch_scatter_t(phj_handle* h, const PassParams& pp, uint32_t grid) {
    if (h->use_match) return launch_scatter_tb<BITS, HASH, POW2, TMA, false>(h, pp, grid);
    return launch_scatter_tb<BITS, HASH, POW2, TMA, true>(h, pp, grid);
}

template <int BITS, int HASH, bool POW2>
cudaError_t launch_hist_lanes_t(phj_handle* h, const PassParams& pp, uint32_t grid) {
    constexpr size_t smem = HistLanesSmem<BITS, PHJ_HIST_TPB>::total;
    auto kern = radix_histogram_lanes<BITS, HASH, POW2, PHJ_HIST_TPB, PHJ_HIST_IPT>;
    static bool configured[16] = {};
    if (!configured[h->device & 15]) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        configured[h->device & 15] = true;
    }
    kern<<<grid, PHJ_HIST_TPB, smem, h->stream>>>(pp);
    return cudaSuccess;
}

cudaError_t launch_hist_full(phj_handle* h, const PassParams& pp, uint32_t grid) {
    switch (h->cfg.hash) {
        case PHJ_HASH_MURMUR3:
            radix_histogram_full<kMurmur3, PHJ_HIST_TPB, PHJ_HIST_IPT><<<grid, PHJ_HIST_TPB, 0, h->stream>>>(pp);
            break;
        case PHJ_HASH_CITY:
            radix_histogram_full<kCity, PHJ_HIST_TPB, PHJ_HIST_IPT><<<grid, PHJ_HIST_TPB, 0, h->stream>>>(pp);
            break;
        default:
            radix_histogram_full<kXXH3, PHJ_HIST_TPB, PHJ_HIST_IPT><<<grid, PHJ_HIST_TPB, 0, h->stream>>>(pp);
            break;
    }
    return cudaGetLastError();
}

template <int BITS, int HASH, bool POW2>
cudaError_t launch_hist_lanes8_t(phj_handle* h, const PassParams& pp, uint32_t grid) {
    constexpr int kTpb = 256;  // 8 warps x 8 KB of byte counters: three CTAs per SM at 256 digits
    constexpr size_t smem = HistLanes8Smem<BITS, kTpb>::total;
    auto kern = radix_histogram_lanes8<BITS, HASH, POW2, kTpb, PHJ_HIST_IPT>;
    static bool configured[16] = {};
    if (!configured[h->device & 15]) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        configured[h->device & 15] = true;
    }
    kern<<<grid, kTpb, smem, h->stream>>>(pp);
    return cudaSuccess;
}

template <int BITS, int HASH>
cudaError_t launch_pass_hp(phj_handle* h, bool scatter, const PassParams& pp, uint32_t grid) {
    const bool tma = !(h->cfg.flags & PHJ_FLAG_NO_TMA_STORE);
    if (BITS > 6 && h->use_lanes && !scatter)  // byte-wide lane-private counters for 128 / 256 digits
        return h->pow2 ? launch_hist_lanes8_t<8, HASH, true>(h, pp, grid)
                       : launch_hist_lanes8_t<8, HASH, false>(h, pp, grid);
    if (BITS <= 6 && h->use_lanes && !scatter)  // lane-private counters: histogram at the HBM roofline
        return h->pow2 ? launch_hist_lanes_t<6, HASH, true>(h, pp, grid)
                       : launch_hist_lanes_t<6, HASH, false>(h, pp, grid);
    if (!scatter) {
        if (h->pow2) launch_hist_t<BITS, HASH, true>(h, pp, grid);
        else launch_hist_t<BITS, HASH, false>(h, pp, grid);
        return cudaSuccess;
    }
    if (h->pow2)
        return tma ? launch_scatter_t<BITS, HASH, true, true>(h, pp, grid)
                   : launch_scatter_t<BITS, HASH, true, false>(h, pp, grid);
    return tma ? launch_scatter_t<BITS, HASH, false, true>(h, pp, grid)
               : launch_scatter_t<BITS, HASH, false, false>(h, pp, grid);
}

template <int BITS>
cudaError_t launch_pass_b(phj_handle* h, bool scatter, const PassParams& pp, uint32_t grid) {
    switch (h->cfg.hash) {
        case PHJ_HASH_MURMUR3: return launch_pass_hp<BITS, kMurmur3>(h, scatter, pp, grid);
        case PHJ_HASH_CITY: return launch_pass_hp<BITS, kCity>(h, scatter, pp, grid);
        default: return launch_pass_hp<BITS, kXXH3>(h, scatter, pp, grid);
    }
}

// The multi-GPU split scatter: tiles of 8192 tuples (1024 threads, one CTA per SM). It is bound by
// NVLink, not by HBM, and NVLink efficiency grows with the length of the per-digit runs a tile
// yields (measured at 2 GPUs, 128 digits: exchange 3.95 / 3.32 / 3.02 ms for 2048 / 4096 / 8192).
template <int BITS, int HASH, bool POW2 = true>
cudaError_t launch_split_scatter_t(phj_handle* h, const PassParams& pp, uint32_t grid) {
    constexpr int kTpb = 1024, kIpt = 8;
    using L = ScatterSmem<BITS, kTpb, kIpt>;
    auto kern = radix_scatter<BITS, HASH, POW2, kTpb, kIpt, true, true, false, 1>;
    static bool configured[16] = {};
    if (!configured[h->device & 15]) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L::total_tma);
        if (e != cudaSuccess) return e;
        configured[h->device & 15] = true;
    }
    kern<<<grid, kTpb, L::total_tma, h->stream>>>(pp);
    return cudaSuccess;
}

cudaError_t launch_split_scatter(phj_handle* h, int bits, const PassParams& pp, uint32_t grid) {
    const bool wide = bits > 6;
    if (!h->pow2) {  // ranks x local partitions is not a power of two (3, 5, 6, 7 ... GPUs): digit = hash % digits
        switch (h->cfg.hash) {
            case PHJ_HASH_MURMUR3:
                return wide ? launch_split_scatter_t<8, kMurmur3, false>(h, pp, grid)
                            : launch_split_scatter_t<6, kMurmur3, false>(h, pp, grid);
            case PHJ_HASH_CITY:
                return wide ? launch_split_scatter_t<8, kCity, false>(h, pp, grid)
                            : launch_split_scatter_t<6, kCity, false>(h, pp, grid);
            default:
                return wide ? launch_split_scatter_t<8, kXXH3, false>(h, pp, grid)
                            : launch_split_scatter_t<6, kXXH3, false>(h, pp, grid);
        }
    }
    switch (h->cfg.hash) {
        case PHJ_HASH_MURMUR3:
            return wide ? launch_split_scatter_t<8, kMurmur3>(h, pp, grid) : launch_split_scatter_t<6, kMurmur3>(h, pp, grid);
        case PHJ_HASH_CITY:
            return wide ? launch_split_scatter_t<8, kCity>(h, pp, grid) : launch_split_scatter_t<6, kCity>(h, pp, grid);
        default:
            return wide ? launch_split_scatter_t<8, kXXH3>(h, pp, grid) : launch_split_scatter_t<6, kXXH3>(h, pp, grid);
    }
}

// A pass with `bits` digit bits runs on the smallest instantiated shape that holds them.
cudaError_t launch_pass(phj_handle* h, bool scatter, int bits, const PassParams& pp, uint32_t grid) {
    if (bits <= 6) return launch_pass_b<6>(h, scatter, pp, grid);
    return launch_pass_b<8>(h, scatter, pp, grid);
}

cudaError_t launch_join(phj_handle* h, const JoinParams& jp, uint32_t grid, size_t smem) {
    auto kern = join_partitions<PHJ_JOIN_TPB, PHJ_JOIN_BUCKET>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    kern<<<grid, PHJ_JOIN_TPB, smem, h->stream>>>(jp);
    return cudaSuccess;
}

template <int HASH>
void launch_gt_t(phj_handle* h, bool build, const GtParams& gp, uint32_t grid) {
    if (!build && (h->cfg.flags & PHJ_FLAG_COOP_PROBE)) {  // A/B: four lanes per probe
        if (h->pow2) gt_probe_coop<HASH, true><<<grid, 256, 0, h->stream>>>(gp);
        else gt_probe_coop<HASH, false><<<grid, 256, 0, h->stream>>>(gp);
        return;
    }
    if (h->pow2) {
        if (build) gt_build<HASH, true><<<grid, 256, 0, h->stream>>>(gp);
        else gt_probe<HASH, true><<<grid, 256, 0, h->stream>>>(gp);
    } else {
        if (build) gt_build<HASH, false><<<grid, 256, 0, h->stream>>>(gp);
        else gt_probe<HASH, false><<<grid, 256, 0, h->stream>>>(gp);
    }
}

void launch_gt(phj_handle* h, bool build, const GtParams& gp, uint32_t grid) {
    switch (h->cfg.hash) {
        case PHJ_HASH_MURMUR3: launch_gt_t<kMurmur3>(h, build, gp, grid); break;
        case PHJ_HASH_CITY: launch_gt_t<kCity>(h, build, gp, grid); break;
        default: launch_gt_t<kXXH3>(h, build, gp, grid); break;
    }
}

// ---- planning -----------------------------------------------------------------------------------
// Fan-out plan from config + |R|: the reference exposes one knob, NumberOfPartitions
// (src/RadixCluster/Configuration.hpp:8); a power of two is split most-significant-digit first
// into <= 2 passes of <= 8 bits so that the final order is the reference's partition order.
int plan_radix(phj_handle* h) {
    const phj_config& c = h->cfg;
    uint64_t P = c.partitions;
    if (P == 0) {
        // engine default: build partitions of ~2.4 K keys (a 4 K-slot shared-memory table)
        uint64_t want = (h->n[0] + 2399) / 2400;
        int b = ilog2_ceil(std::max<uint64_t>(want, 1));
        b = std::min(b, 2 * kMaxBitsPerPass);
        P = 1ull << b;
    }
    h->P = P;
    h->pow2 = (P & (P - 1)) == 0;
    h->bits_total = ilog2_ceil(P);
    if (h->bits_total > 2 * kMaxBitsPerPass)
        return fail(PHJ_ERR_INVALID, "partitions=%llu exceeds the supported fan-out of %llu",
                    (unsigned long long)P, 1ull << (2 * kMaxBitsPerPass));
    int b1 = (int)c.radix_bits[0], b2 = (int)c.radix_bits[1];
    if (c.algo == PHJ_ALGO_SHARD_SPLIT) {
        b1 = h->bits_total;
        b2 = 0;
    } else if (b1 == 0 && b2 == 0) {
        if (h->bits_total <= 7) {
            b1 = h->bits_total;
            b2 = 0;
        } else {
            b1 = (h->bits_total + 1) / 2;
            b2 = h->bits_total - b1;
        }
    } else {
        if (b1 + b2 != h->bits_total || b1 > kMaxBitsPerPass || b2 > kMaxBitsPerPass || b1 < 0 ||
            b2 < 0 || (b1 == 0 && h->bits_total > 0))
            return fail(PHJ_ERR_INVALID,
                        "radix_bits {%d,%d} must each be <= %d and sum to log2(partitions) = %d", b1,
                        b2, kMaxBitsPerPass, h->bits_total);
    }
    h->b1 = b1;
    h->b2 = b2;
    h->d2 = 1u << b2;
    h->d1 = (uint32_t)((P + h->d2 - 1) >> b2);  // pass-1 digits actually populated
    // A pre-partitioned bind may own any contiguous range of the pass-1 digits (the multi-GPU path
    // balances digit ranges over ranks by their tuple counts): pass 2 and the join only need the count.
    if (h->prepart_parents) h->d1 = h->prepart_parents;
    h->nparts = (uint64_t)h->d1 * h->d2;
    return PHJ_OK;
}

DigitFn digit_fn(const phj_handle* h, int pass) {
    DigitFn f;
    f.pmask = h->pow2 ? h->P - 1 : ~0ull;
    f.modulus = h->P;
    if (h->cfg.algo == PHJ_ALGO_SHARD_SPLIT) {  // owner rank = hash bits [shard_shift, +log2 ranks)
        f.pmask = ~0ull;
        f.shift = h->cfg.shard_shift;
        f.mask = (uint32_t)(h->P - 1);
        if (!h->pow2) {  // digit = hash % P (the kernels' POW2 = false path; shard_shift is 0 then)
            f.shift = 0;
            f.mask = ~0u;
        }
        return f;
    }
    if (pass == 1) {
        f.shift = (uint32_t)h->b2;
        f.mask = (1u << std::max(h->b1, 1)) - 1;  // b1 == 0 only when P == 1 (digit always 0)
        if (h->b1 == 0) f.mask = 0;
    } else {
        f.shift = 0;
        f.mask = (1u << h->b2) - 1;
    }
    return f;
}

uint32_t segments_for(const phj_handle* h, size_t n, int tile) {
    if (n == 0) return 0;
    // ~8 segments per SM for the big relation; never shorter than 4 tiles
    uint64_t target = (uint64_t)h->sm_count * PHJ_SEGS_PER_SM;
    uint64_t by_len = (n + (uint64_t)tile * 4 - 1) / ((uint64_t)tile * 4);
    return (uint32_t)std::max<uint64_t>(1, std::min(target, by_len));
}

// The last boundary of every relation (= n) is never written by a kernel.
int upload_tail_bounds(phj_handle* h) {
    for (int rel = 0; rel < 2; ++rel) {
        h->h_out[4 + rel] = h->n[rel];  // pinned staging
        PHJ_CUDA(cudaMemcpyAsync(h->d_bounds1[rel] + h->d1, h->h_out + 4 + rel, 8, cudaMemcpyHostToDevice, h->stream));
        PHJ_CUDA(cudaMemcpyAsync(h->d_bounds2[rel] + h->nparts, h->h_out + 4 + rel, 8, cudaMemcpyHostToDevice, h->stream));
    }
    PHJ_CUDA(cudaStreamSynchronize(h->stream));
    return PHJ_OK;
}

int build_plan(phj_handle* h) {
    int rc;
    if (h->cfg.algo == PHJ_ALGO_NO_PARTITIONING) {
        h->P = 1;
        h->pow2 = true;
        h->nparts = 1;
        for (int rel = 0; rel < 2; ++rel) h->plan_n[rel] = h->segs_n[rel] = h->n[rel];
    } else {
        if ((rc = plan_radix(h)) != PHJ_OK) return rc;
        // ---- pass 1 segments (host-built: sizes are known) ----
        std::vector<Segment> segs;
        uint32_t nseg[2];
        uint32_t cnt_base = 0;
        uint32_t nseg_plan[2];
        for (int rel = 0; rel < 2; ++rel) {
            // sized for max(n, reserve): a later bind of fewer pre-partitioned tuples keeps the plan
            const size_t pn = std::max<size_t>(h->n[rel], rel == 0 ? h->cfg.reserve_build : h->cfg.reserve_probe);
            h->plan_n[rel] = pn;
            nseg_plan[rel] = segments_for(h, pn, kScatTile);
            uint64_t len = nseg_plan[rel] ? (pn + nseg_plan[rel] - 1) / nseg_plan[rel] : 0;
            len = ((len + kScatTile - 1) / kScatTile) * kScatTile;
            nseg_plan[rel] = len ? (uint32_t)((pn + len - 1) / len) : 0;
            nseg[rel] = len ? (uint32_t)((h->n[rel] + len - 1) / len) : 0;
            h->seg_len[rel] = len ? len : kScatTile;
            // radix_histogram_lanes keeps lane-private 16-bit counters for a whole segment: a thread must see
            // fewer than 65536 tuples per segment whatever their digits are
            if (len / PHJ_HIST_TPB >= 65536)
                return fail(PHJ_ERR_INVALID, "internal: segments of %llu tuples overflow the histogram's 16-bit lane counters",
                            (unsigned long long)len);
            h->nseg1_rel[rel] = nseg[rel];
            h->cnt_base1_rel[rel] = cnt_base;
            for (uint32_t s = 0; s < nseg[rel]; ++s) {
                Segment sg;
                sg.begin = (uint64_t)s * len;
                sg.end = std::min<uint64_t>(h->n[rel], sg.begin + len);
                sg.cnt_index = cnt_base + s;
                sg.cnt_stride = nseg[rel];
                sg.rel = rel;
                sg.parent_first = (s == 0 ? 0x80000000u : 0u);
                segs.push_back(sg);
            }
            cnt_base += nseg[rel] * h->d1;
        }
        h->nsegs1 = (uint32_t)segs.size();
        uint32_t ncounts1 = cnt_base;
        // row chunks of the multi-GPU split: whole segments; the build relation goes with chunk 0
        h->nchunks = h->cfg.algo == PHJ_ALGO_SHARD_SPLIT ? std::max<uint32_t>(1, h->cfg.split_chunks) : 1;
        for (uint32_t c = 0; c <= h->nchunks; ++c) {
            h->chunk_first_seg[0][c] = c == 0 ? 0 : nseg[0];
            h->chunk_first_seg[1][c] = (uint32_t)((uint64_t)nseg[1] * c / h->nchunks);
        }
        h->segs_n[0] = h->n[0];
        h->segs_n[1] = h->n[1];
        if ((rc = dev_reserve(&h->d_segs1, &h->cap_segs1, segs.size())) != PHJ_OK) return rc;
        if (!segs.empty())
            PHJ_CUDA(cudaMemcpyAsync(h->d_segs1, segs.data(), segs.size() * sizeof(Segment),
                                     cudaMemcpyHostToDevice, h->stream));
        // ---- pass 2 capacity ----
        size_t ncounts_max = ncounts1;
        h->fuse2 = false;
        if (h->b2 > 0) {
            const uint32_t div = 1;
            for (int rel = 0; rel < 2; ++rel) {
                h->seg_len2[rel] = ((h->seg_len[rel] / div + kScatTile - 1) / kScatTile) * kScatTile;
                h->target_segs2[rel] = std::max<uint32_t>(1, nseg_plan[rel]) * div + div;
            }
            h->max_segs2 = h->target_segs2[0] + h->target_segs2[1] + 2 * h->d1;
            // The pass-1 scatter also counts every tuple into its pass-2 segment, so pass 2 needs no
            // histogram read of its own: 4.39 -> 4.28 ms at 10 M x 200 M (uniform), 4.40 -> 4.33 ms at
            // Zipf 1.25. On by default for two 6-bit passes since the counters are packed 16-bit (two
            // CTAs per SM); with 32-bit counters only one CTA fit and it was slower than the extra read.
            h->fuse2 = h->b1 <= 6 && h->b2 <= 6 && !(h->cfg.flags & PHJ_FLAG_NO_FUSE_HIST2) && div == 1 &&
                       h->cfg.algo == PHJ_ALGO_RADIX_PARTITIONING;
            // Both passes' histograms from ONE read of the input (radix_histogram_full, K1d): the
            // pass-1 scatter then runs without the fused counting (1.57 -> 1.29 ms) for a histogram
            // that takes 0.53 instead of 0.49 ms. PHJ_FLAG_NO_HIST12 falls back to the fused histogram.
            {
                h->hist12 = !(h->cfg.flags & PHJ_FLAG_NO_HIST12) && h->fuse2 && h->pow2 && h->d1 <= kFullD1 &&
                            h->d2 <= kFullD2;
                if (h->hist12 &&
                    (rc = dev_reserve(&h->d_hist12, &h->cap_hist12,
                                      (size_t)std::max<uint32_t>(h->nsegs1, 1) * kFullD1 * kFullD2)) != PHJ_OK)
                    return rc;
            }
            if (h->d1 > h->cap_parents2 || !h->d_parents2[0]) {
                for (int rel = 0; rel < 2; ++rel) {
                    if (h->d_parents2[rel]) cudaFree(h->d_parents2[rel]);
                    PHJ_CUDA(cudaMalloc(&h->d_parents2[rel], (size_t)h->d1 * sizeof(Parent2)));
                }
                h->cap_parents2 = h->d1;
            }
            if ((rc = dev_reserve(&h->d_segs2, &h->cap_segs2, h->max_segs2)) != PHJ_OK) return rc;
            ncounts_max = std::max<size_t>(ncounts_max, (size_t)h->max_segs2 * h->d2);
        }
        if (ncounts_max > (size_t)kScanChunk * 1024)
            return fail(PHJ_ERR_INVALID, "internal: %zu counters exceed the scan capacity", ncounts_max);
        if ((rc = dev_reserve(&h->d_counts, &h->cap_counts, ncounts_max)) != PHJ_OK) return rc;
        if ((rc = dev_reserve(&h->d_cursors, &h->cap_cursors, ncounts_max)) != PHJ_OK) return rc;
        if (!h->d_scan_state) {
            PHJ_CUDA(cudaMalloc(&h->d_scan_state, sizeof(ScanState)));
            PHJ_CUDA(cudaMemset(h->d_scan_state, 0, sizeof(ScanState)));
        }
        // ---- boundaries ----
        const size_t nb1 = (size_t)h->d1 + 1, nb2 = (size_t)h->nparts + 1;
        if (nb1 > h->cap_bounds1 || !h->d_bounds1[0]) {
            for (int rel = 0; rel < 2; ++rel) {
                if (h->d_bounds1[rel]) cudaFree(h->d_bounds1[rel]);
                PHJ_CUDA(cudaMalloc(&h->d_bounds1[rel], nb1 * 8));
            }
            h->cap_bounds1 = nb1;
        }
        if (nb2 > h->cap_bounds2 || !h->d_bounds2[0]) {
            for (int rel = 0; rel < 2; ++rel) {
                if (h->d_bounds2[rel]) cudaFree(h->d_bounds2[rel]);
                PHJ_CUDA(cudaMalloc(&h->d_bounds2[rel], nb2 * 8));
            }
            h->cap_bounds2 = nb2;
        }
        for (int rel = 0; rel < 2; ++rel) {
            // a relation without tuples launches no segment: its boundaries are all zero. The last
            // boundary (= n) is never written by a kernel.
            PHJ_CUDA(cudaMemsetAsync(h->d_bounds1[rel], 0, nb1 * 8, h->stream));
            PHJ_CUDA(cudaMemsetAsync(h->d_bounds2[rel], 0, nb2 * 8, h->stream));
        }
        if ((rc = upload_tail_bounds(h)) != PHJ_OK) return rc;
        uint32_t sc[kNumScalars] = {};
        sc[kNsegs1] = h->nsegs1;
        sc[kNcounts1] = ncounts1;
        PHJ_CUDA(cudaMemcpyAsync(h->d_scalars, sc, sizeof(sc), cudaMemcpyHostToDevice, h->stream));
        PHJ_CUDA(cudaStreamSynchronize(h->stream));

        // ---- partition buffers: sized here, allocated by ensure_buffers() for the passes a join
        // really runs (a pre-partitioned join needs no pass-1 output, a remote split none at all) ----
        for (int rel = 0; rel < 2; ++rel) {
            if (h->n[rel] > h->cap_buf[rel]) {
                if (h->d_buf_a[rel]) cudaFree(h->d_buf_a[rel]);
                if (h->d_buf_b[rel]) cudaFree(h->d_buf_b[rel]);
                h->d_buf_a[rel] = h->d_buf_b[rel] = nullptr;
                h->cap_buf[rel] = std::max<size_t>(std::max<size_t>(h->n[rel], 1),
                                                   rel == 0 ? h->cfg.reserve_build : h->cfg.reserve_probe);
            }
        }
        // ---- join geometry ----
        // Table of 32-byte buckets in shared memory: aim for a load factor of ~0.3 (nearly every
        // probe ends in its home bucket) while the table stays <= 64 KB (3 CTAs per SM); beyond that
        // accept ~0.6, and beyond 128 KB the partition goes through the global table.
        uint64_t mean = std::max<uint64_t>(1, h->plan_n[0] / std::max<uint64_t>(1, h->P));
        uint32_t slots = 1024;
        while (slots < 8192 && (double)slots < 3.2 * (double)mean) slots <<= 1;
        while (slots < 16384 && (double)slots < 1.6 * (double)mean) slots <<= 1;
        h->join_slots = slots;
        h->join_max_keys = slots / 4 * 3;
        const size_t smem = (size_t)slots * 8;
        uint32_t resident = (uint32_t)std::max<size_t>(1, std::min<size_t>((h->smem_optin) / (smem + 1024), 2048 / PHJ_JOIN_TPB));
        h->join_grid = (uint32_t)h->sm_count * resident * PHJ_JOIN_WAVES;
        {
            // PHJ_FLAG_L2_TABLES: L2-resident partition tables instead of shared-memory ones. A region of
            // 32-byte buckets per partition, load <= 0.4 at the mean partition size; partitions beyond 75 % of
            // a region's capacity take the oversize path like partitions too large for shared memory.
            h->l2join = (h->cfg.flags & PHJ_FLAG_L2_TABLES) && h->cfg.algo == PHJ_ALGO_RADIX_PARTITIONING;
            if (h->l2join) {
                uint32_t rb = 64;
                while ((uint64_t)rb * 8 < mean * 5) rb <<= 1;  // 4 rb >= 2.5 x mean
                h->pt_region_buckets = rb;
                h->join_max_keys = rb * 3;
                if ((rc = dev_reserve(&h->d_pt, &h->cap_pt, (size_t)h->nparts * rb * 4)) != PHJ_OK) return rc;
            }
        }
        size_t cap = h->cap_cta_times;
        if ((rc = dev_reserve(&h->d_cta_times, &cap, (size_t)h->join_grid * 2)) != PHJ_OK) return rc;
        if (cap != h->cap_cta_times) {
            if (h->h_cta_times) cudaFreeHost(h->h_cta_times);
            PHJ_CUDA(cudaMallocHost(&h->h_cta_times, cap * 8));
            h->cap_cta_times = cap;
        }
    }
    // ---- bucket-chained table: ceil(0.25 n) heads rounded to a power of two, one preallocated
    // bucket per head + ceil(n / 3) overflow buckets (SeparateChaining.hpp:150-181, SeparateChaining.cpp) ----
    if (h->cfg.algo == PHJ_ALGO_NO_PARTITIONING && (h->cfg.flags & PHJ_FLAG_CHAINED_TABLE)) {
        const uint64_t n0 = std::max<size_t>(h->n[0], 1);
        uint64_t heads = 64;
        while (heads < (n0 + 3) / 4) heads <<= 1;
        const uint64_t pool = heads + (n0 + 2) / 3 + 1024;
        if (pool >= 0xfffffff0ull) return fail(PHJ_ERR_INVALID, "build relation too large for the chained table");
        if (heads > h->ct_nheads || pool > h->ct_pool) {
            if (h->d_ct_heads) cudaFree(h->d_ct_heads);
            if (h->d_ct_buckets) cudaFree(h->d_ct_buckets);
            h->d_ct_heads = nullptr;
            h->d_ct_buckets = nullptr;
            PHJ_CUDA(cudaMalloc(&h->d_ct_heads, heads * 4));
            PHJ_CUDA(cudaMalloc(&h->d_ct_buckets, pool * sizeof(ChainBucket)));
        }
        h->ct_nheads = (uint32_t)heads;
        h->ct_pool = (uint32_t)pool;
        return PHJ_OK;
    }
    // ---- global table: always for NPJ; for radix it is allocated on first need ----
    if (h->cfg.algo == PHJ_ALGO_NO_PARTITIONING) {
        uint64_t want = (uint64_t)std::ceil((double)std::max<size_t>(h->n[0], 1) / 2.8);
        uint64_t buckets = 64;
        while (buckets < want) buckets <<= 1;
        if (buckets > h->gt_buckets) {
            if (h->d_gt) cudaFree(h->d_gt);
            h->d_gt = nullptr;
            PHJ_CUDA(cudaMalloc(&h->d_gt, buckets * 32));
        }
        h->gt_buckets = buckets;
    }
    return PHJ_OK;
}

// Pass-1 / pass-2 output buffers, allocated on first need (phj_upload / phj_bind_* call this for
// the plan at hand, so phj_join itself allocates nothing).
int ensure_buffers(phj_handle* h, bool need_a, bool need_b) {
    for (int rel = 0; rel < 2; ++rel) {
        const size_t want = std::max<size_t>(h->cap_buf[rel], 1);
        if (need_a && !h->d_buf_a[rel]) PHJ_CUDA(cudaMalloc(&h->d_buf_a[rel], want * 16));
        if (need_b && !h->d_buf_b[rel]) PHJ_CUDA(cudaMalloc(&h->d_buf_b[rel], want * 16));
    }
    return PHJ_OK;
}

int ensure_gt_for_fallback(phj_handle* h) {
    uint64_t want = (uint64_t)std::ceil((double)std::max<size_t>(h->n[0], 1) / 2.8);
    uint64_t buckets = 64;
    while (buckets < want) buckets <<= 1;
    if (!h->d_gt || buckets > h->gt_buckets) {
        if (h->d_gt) cudaFree(h->d_gt);
        h->d_gt = nullptr;
        PHJ_CUDA(cudaMalloc(&h->d_gt, buckets * 32));
        h->gt_buckets = buckets;
    }
    return PHJ_OK;
}

uint32_t table_hash_shift(const phj_handle* h) {
    // the table takes hash bits above the ones the partitioning consumed
    if (h->cfg.algo == PHJ_ALGO_NO_PARTITIONING) return 0u;
    int bits = h->bits_total;
    // Relations that arrived partitioned by a WIDER pass-1 digit than this handle's own (the owner-rank bits of the
    // multi-GPU split sit right above it) are nearly constant in those bits too: skip them, or only 1 / world of the
    // buckets would be home buckets and the rest would fill by overflow alone.
    if (h->prepart && h->prepart_space > ((uint64_t)1 << h->b1)) bits += ilog2_ceil(h->prepart_space) - h->b1;
    return (uint32_t)std::min(bits, 24);
}

// ---- the joins ----------------------------------------------------------------------------------
int run_scan(phj_handle* h, int ncounts_scalar, size_t ncounts_max) {
    uint32_t chunks = (uint32_t)std::max<size_t>(1, (ncounts_max + kScanChunk - 1) / kScanChunk);
    {
        KernelScope ks(h, "scan");
        scan_lookback<<<chunks, kScanTpb, 0, h->stream>>>(h->d_counts, h->d_scalars + ncounts_scalar, h->d_cursors,
                                                          h->d_scan_state, ++h->scan_epoch);
    }
    return PHJ_OK;
}

int run_gt(phj_handle* h, bool select, const ulonglong2* build, const ulonglong2* probe) {
    const HashParams hp = make_hash_params(h->cfg.hash, h->cfg.hash_seed);
    GtParams gp{};
    gp.table = h->d_gt;
    gp.bucket_mask = h->gt_buckets - 1;
    gp.hash_shift = table_hash_shift(h);
    gp.select = select ? 1 : 0;
    gp.bounds_build = select ? h->d_bounds2[0] : nullptr;
    gp.max_keys = h->join_max_keys;
    gp.part_fn.pmask = h->pow2 ? h->P - 1 : ~0ull;
    gp.part_fn.modulus = h->P;
    gp.part_fn.shift = 0;
    gp.part_fn.mask = ~0u;
    if (select && h->prepart) {
        const uint64_t space = h->prepart_space ? h->prepart_space : ((uint64_t)1 << h->b1);
        gp.pre_space_mask = space - 1 ? space - 1 : 0;
        gp.pre_first = h->prepart_first;
        gp.pre_b2 = (uint32_t)h->b2;
    }
    if (select && h->pow2 && h->bits_total > 0 && !(h->prepart && h->prepart_parents != (1u << h->b1))) {
        // partition-local table regions (at most 1024 of them, never smaller than 1024 buckets)
        uint32_t rb = (uint32_t)std::min(h->bits_total, 10);
        const uint32_t bucket_bits = (uint32_t)ilog2_ceil(h->gt_buckets);
        if (bucket_bits < rb + 10) rb = bucket_bits > 10 ? bucket_bits - 10 : 0;
        gp.region_bits = rb;
        gp.region_shift = (uint32_t)h->bits_total - rb;
        gp.local_mask = (h->gt_buckets >> rb) - 1;
    }
    gp.hp = hp;
    gp.flags = h->d_scalars + kGtFlags;
    gp.matches = h->d_matches;
    const uint32_t grid = (uint32_t)h->sm_count * 8;
    {
        KernelScope ks(h, "gt_clear");
        PHJ_CUDA(cudaMemsetAsync(h->d_scalars + kGtFlags, 0, 4, h->stream));
        gt_clear<<<grid, 256, 0, h->stream>>>(h->d_gt, h->gt_buckets * 4);
    }
    {
        KernelScope ks(h, "gt_build");
        gp.rel = build;
        gp.n = h->n[0];
        launch_gt(h, true, gp, grid);
    }
    PHJ_CUDA(cudaEventRecord(h->ev[2], h->stream));
    {
        KernelScope ks(h, "gt_probe");
        gp.rel = probe;
        gp.n = h->n[1];
        launch_gt(h, false, gp, grid);
    }
    return PHJ_OK;
}

template <int HASH>
void launch_ct_t(phj_handle* h, bool build, const CtParams& cp, uint32_t grid) {
    if (build) ct_build<HASH><<<grid, 256, 0, h->stream>>>(cp);
    else ct_probe<HASH><<<grid, 256, 0, h->stream>>>(cp);
}

// No-partitioning join through the bucket-chained table.
int run_ct(phj_handle* h, const ulonglong2* build, const ulonglong2* probe) {
    CtParams cp{};
    cp.heads = h->d_ct_heads;
    cp.buckets = h->d_ct_buckets;
    cp.nheads_mask = h->ct_nheads - 1;
    cp.hash_shift = 0;
    cp.pool = h->ct_pool;
    cp.cursor = h->d_scalars + kCtCursor;
    cp.overflow = h->d_scalars + kGtFlags;
    cp.hp = make_hash_params(h->cfg.hash, h->cfg.hash_seed);
    cp.matches = h->d_matches;
    const uint32_t grid = (uint32_t)h->sm_count * 8;
    auto launch = [&](bool is_build) {
        switch (h->cfg.hash) {
            case PHJ_HASH_MURMUR3: launch_ct_t<kMurmur3>(h, is_build, cp, grid); break;
            case PHJ_HASH_CITY: launch_ct_t<kCity>(h, is_build, cp, grid); break;
            default: launch_ct_t<kXXH3>(h, is_build, cp, grid); break;
        }
    };
    {
        KernelScope ks(h, "ct_init");
        ct_init<<<grid, 256, 0, h->stream>>>(cp);
    }
    {
        KernelScope ks(h, "ct_build");
        cp.rel = build;
        cp.n = h->n[0];
        launch(true);
    }
    PHJ_CUDA(cudaEventRecord(h->ev[2], h->stream));
    {
        KernelScope ks(h, "ct_probe");
        cp.rel = probe;
        cp.n = h->n[1];
        launch(false);
    }
    return PHJ_OK;
}

float ev_ms(cudaEvent_t a, cudaEvent_t b) {
    float ms = 0;
    cudaEventElapsedTime(&ms, a, b);
    return ms;
}

int join_no_partitioning(phj_handle* h, phj_result* out) {
    // Deviation: the reference's table constructor throws for an empty build side
    // (src/HashTables/LinearProbing.hpp:106-110); here an empty R simply joins to 0.
    PHJ_CUDA(cudaEventRecord(h->ev[0], h->stream));
    PHJ_CUDA(cudaMemsetAsync(h->d_matches, 0, 8, h->stream));
    const bool chained = (h->cfg.flags & PHJ_FLAG_CHAINED_TABLE) != 0;
    int rc = chained ? run_ct(h, h->d_in[0], h->d_in[1]) : run_gt(h, false, h->d_in[0], h->d_in[1]);
    if (rc != PHJ_OK) return rc;
    PHJ_CUDA(cudaEventRecord(h->ev[3], h->stream));
    PHJ_CUDA(cudaMemcpyAsync(h->h_out, h->d_matches, 8, cudaMemcpyDeviceToHost, h->stream));
    PHJ_CUDA(cudaMemcpyAsync(h->h_out + 1, h->d_scalars + kGtFlags, 4, cudaMemcpyDeviceToHost, h->stream));
    PHJ_CUDA(cudaEventRecord(h->ev[4], h->stream));
    PHJ_CUDA(cudaStreamSynchronize(h->stream));
    PHJ_CUDA(cudaGetLastError());
    if (chained && (h->h_out[1] & 0xffffffffu))
        return fail(PHJ_ERR_NOMEM, "BucketAllocator exceeded its limit.");  // SeparateChaining.hpp:114-118
    out->matches = h->h_out[0];
    out->partition_ns = 0;
    out->build_ns = (uint64_t)(ev_ms(h->ev[0], h->ev[2]) * 1e6);
    out->probe_ns = (uint64_t)(ev_ms(h->ev[2], h->ev[3]) * 1e6);
    out->join_ns = out->build_ns + out->probe_ns;
    out->total_ns = (uint64_t)(ev_ms(h->ev[0], h->ev[4]) * 1e6);
    out->passes = 0;
    out->partitions = 1;
    out->d2h_bytes = 8;
    // 16 B/tuple streamed + one 32-byte table sector per build insert and per probe
    out->hbm_bytes_alg = 48ull * (h->n[0] + h->n[1]);
    return PHJ_OK;
}

int join_radix(phj_handle* h, phj_result* out) {
    const HashParams hp = make_hash_params(h->cfg.hash, h->cfg.hash_seed);
    const bool two = h->b2 > 0;
    {
        int rc = ensure_buffers(h, !h->prepart, two);  // no-op unless the plan changed since the bind
        if (rc != PHJ_OK) return rc;
    }
    PHJ_CUDA(cudaEventRecord(h->ev[0], h->stream));

    // ---- pass 1 ----
    PassParams p1{};
    for (int rel = 0; rel < 2; ++rel) {
        p1.in[rel] = h->d_in[rel];
        p1.out[rel] = h->d_buf_a[rel];
        p1.bounds[rel] = two ? h->d_bounds1[rel] : h->d_bounds2[rel];
    }
    p1.segs = h->d_segs1;
    p1.nsegs = h->d_scalars + kNsegs1;
    p1.counts = h->d_counts;
    p1.cursors = h->d_cursors;
    p1.cursor_bias[0] = 0;
    p1.cursor_bias[1] = h->n[0];
    p1.bounds_stride = h->d1;
    p1.ndigits = h->d1;
    p1.hp = hp;
    p1.df = digit_fn(h, 1);
    // pass-2 bookkeeping (needed before scatter 1 when the pass-2 histogram is fused into it)
    Plan2Params pl{};
    PassParams p2{};
    if (two) {
        for (int rel = 0; rel < 2; ++rel) {
            pl.bounds1[rel] = h->d_bounds1[rel];
            pl.parents2[rel] = h->d_parents2[rel];
            pl.cnt_base1[rel] = h->cnt_base1_rel[rel];
            pl.nseg1[rel] = h->nseg1_rel[rel];
            pl.seg_len[rel] = h->seg_len2[rel];
            pl.n[rel] = h->n[rel];
            pl.target_segs[rel] = h->target_segs2[rel];
            p2.in[rel] = h->d_buf_a[rel];
            p2.out[rel] = h->d_buf_b[rel];
            p2.bounds[rel] = h->d_bounds2[rel];
        }
        pl.cursors = h->d_cursors;
        pl.bias[0] = 0;
        pl.bias[1] = h->n[0];
        pl.d1 = h->d1;
        pl.d2 = h->d2;
        pl.tile = kScatTile;
        pl.segs = h->d_segs2;
        pl.nsegs = h->d_scalars + kNsegs2;
        pl.ncounts = h->d_scalars + kNcounts2;
        pl.max_segs = h->max_segs2;
        pl.truncated = reinterpret_cast<uint32_t*>(h->d_matches + 1) + 1;
        pl.zero_words = h->d_matches;  // the plan kernel also clears the join's result words
        for (int rel = 0; rel < 2; ++rel) pl.bounds2[rel] = h->d_bounds2[rel];
        p2.segs = h->d_segs2;
        p2.nsegs = h->d_scalars + kNsegs2;
        p2.counts = h->d_counts;
        p2.cursors = h->d_cursors;
        p2.cursor_bias[0] = 0;
        p2.cursor_bias[1] = h->n[0];
        p2.bounds_stride = h->d2;
        p2.ndigits = h->d2;
        p2.hp = hp;
        p2.df = digit_fn(h, 2);
    }
    bool results_cleared = false;
    auto run_plan2 = [&]() {
        KernelScope ks(h, "plan_pass2");
        plan_pass2<<<1, 1024, 0, h->stream>>>(pl);
        results_cleared = true;
    };
    const bool hist12 = two && h->hist12 && !h->prepart && h->nsegs1 > 0;
    const bool fuse2 = two && h->fuse2 && !h->prepart && !hist12;
    if (h->prepart) {
        // The relations arrived partitioned by the pass-1 digit (the multi-GPU shuffle was pass 1):
        // their boundaries stand in for the scanned pass-1 cursors, pass 2 reads them in place.
        if (two) {
            pl.cursors = h->d_pre_bounds;
            for (int rel = 0; rel < 2; ++rel) {
                pl.cnt_base1[rel] = rel * (h->d1 + 1);
                pl.nseg1[rel] = 1;
                pl.bias[rel] = 0;
                p2.in[rel] = h->d_in[rel];
            }
        } else {
            for (int rel = 0; rel < 2; ++rel)
                PHJ_CUDA(cudaMemcpyAsync(h->d_bounds2[rel], h->d_pre_bounds + rel * (h->d1 + 1), (h->d1 + 1) * 8,
                                         cudaMemcpyDeviceToDevice, h->stream));
        }
    } else if (h->nsegs1 > 0) {
        {
            KernelScope ks(h, "radix_histogram[1]");
            if (hist12) {
                p1.df2 = digit_fn(h, 2);
                p1.d2 = h->d2;
                p1.hist12 = h->d_hist12;
                PHJ_CUDA(launch_hist_full(h, p1, h->nsegs1));
            } else {
                PHJ_CUDA(launch_pass(h, false, h->b1, p1, h->nsegs1));
            }
        }
        run_scan(h, kNcounts1, (size_t)h->nsegs1 * h->d1);
        if (hist12) {
            // The scanned cursors hold the pass-1 boundaries and every run's start: plan pass 2, align
            // its segments to the runs and sum the (segment, digit pair) counts into pass 2's counters
            // -- all before the pass-1 cursors are overwritten by pass 2's scan.
            run_plan2();
            AlignCounts2Params ap{};
            ap.cursors = h->d_cursors;
            for (int rel = 0; rel < 2; ++rel) {
                ap.cnt_base1[rel] = h->cnt_base1_rel[rel];
                ap.nseg1[rel] = h->nseg1_rel[rel];
                ap.bias[rel] = pl.bias[rel];
                ap.bounds1[rel] = h->d_bounds1[rel];
                ap.parents2[rel] = h->d_parents2[rel];
                ap.seg_len[rel] = h->seg_len2[rel];
            }
            ap.seg_first1[0] = 0;
            ap.seg_first1[1] = h->nseg1_rel[0];
            ap.segs = h->d_segs2;
            ap.nsegs = h->d_scalars + kNsegs2;
            ap.hist12 = h->d_hist12;
            ap.d2 = h->d2;
            ap.counts2 = h->d_counts;
            {
                KernelScope ks(h, "pass2_counts");
                pass2_align_counts<<<(h->max_segs2 + 3) / 4, 256, 0, h->stream>>>(ap);
            }
        } else if (fuse2) {
            // the scanned cursors already hold the pass-1 boundaries: plan pass 2 now, and let the
            // pass-1 scatter count every tuple into its pass-2 segment (no second histogram read)
            run_plan2();
            PHJ_CUDA(cudaMemsetAsync(h->d_counts, 0, (size_t)h->max_segs2 * h->d2 * sizeof(uint32_t), h->stream));
            for (int rel = 0; rel < 2; ++rel) {
                p1.parents2[rel] = h->d_parents2[rel];
                p1.seg_len2[rel] = h->seg_len2[rel];
            }
            p1.counts2 = h->d_counts;
            p1.df2 = digit_fn(h, 2);
            p1.d2 = h->d2;
        }
        {
            KernelScope ks(h, "radix_scatter[1]");
            PHJ_CUDA(launch_pass(h, true, h->b1, p1, h->nsegs1));
        }
    } else if (fuse2) {
        run_plan2();
        PHJ_CUDA(cudaMemsetAsync(h->d_counts, 0, (size_t)h->max_segs2 * h->d2 * sizeof(uint32_t), h->stream));
    }
    // ---- pass 2 ----
    if (two) {
        if (!fuse2 && !hist12) {
            run_plan2();
            KernelScope ks(h, "radix_histogram[2]");
            PHJ_CUDA(launch_pass(h, false, h->b2, p2, h->max_segs2));
        }
        run_scan(h, kNcounts2, (size_t)h->max_segs2 * h->d2);
        {
            KernelScope ks(h, "radix_scatter[2]");
            PHJ_CUDA(launch_pass(h, true, h->b2, p2, h->max_segs2));
        }
    }
    PHJ_CUDA(cudaEventRecord(h->ev[1], h->stream));
    if (h->cfg.algo == PHJ_ALGO_SHARD_SPLIT) {
        PHJ_CUDA(cudaStreamSynchronize(h->stream));
        PHJ_CUDA(cudaGetLastError());
        out->partition_ns = (uint64_t)(ev_ms(h->ev[0], h->ev[1]) * 1e6);
        out->total_ns = out->partition_ns;
        out->passes = 1;
        out->partitions = h->P;
        out->hbm_bytes_alg = 16ull * 3 * (h->n[0] + h->n[1]);  // histogram read + scatter read/write
        h->joined_radix = true;
        return PHJ_OK;
    }

    // ---- build + probe per partition ----
    const ulonglong2* part_build = two ? h->d_buf_b[0] : h->prepart ? h->d_in[0] : h->d_buf_a[0];
    const ulonglong2* part_probe = two ? h->d_buf_b[1] : h->prepart ? h->d_in[1] : h->d_buf_a[1];
    if (!results_cleared) PHJ_CUDA(cudaMemsetAsync(h->d_matches, 0, 16, h->stream));  // else: plan_pass2 did
    uint32_t* d_oversize = reinterpret_cast<uint32_t*>(h->d_matches + 1);
    JoinParams jp{};
    jp.build = part_build;
    jp.probe = part_probe;
    jp.bounds_build = h->d_bounds2[0];
    jp.bounds_probe = h->d_bounds2[1];
    jp.n_probe = h->n[1];
    jp.slice_len = h->n[1] / h->join_grid;
    jp.slice_rem = h->n[1] % h->join_grid;
    jp.npart = (uint32_t)h->nparts;
    jp.bucket_mask = h->join_slots / PHJ_JOIN_BUCKET - 1;
    jp.bucket_shift = 64 - ilog2_ceil(h->join_slots / PHJ_JOIN_BUCKET);
    jp.max_keys = h->join_max_keys;
    jp.table_mul = (h->cfg.table_seed * 0x9E3779B97F4A7C15ULL) | 1ULL;  // odd
    if (h->cfg.table_seed == 0) jp.table_mul = 0xBF58476D1CE4E5B9ULL;
    jp.table_mul |= 1ULL << 32;  // both 32-bit multipliers odd
    jp.shift32 = 32 - ilog2_ceil(h->join_slots / PHJ_JOIN_BUCKET);
    {
        // the partition the free-slot marker (INT64_MIN) belongs to, in the final partition numbering
        const uint64_t hs = hash_key_dyn(h->cfg.hash, kEmptyKey, hp);
        const uint64_t full = h->pow2 ? (hs & (h->P - 1)) : (hs % h->P);
        jp.sentinel_part = (uint32_t)full;
        if (h->prepart) {
            // local parents 0 .. d1 - 1 are digits [first, first + d1) of a pass-1 digit space that may
            // be wider than this handle's own (the owner-rank bits of the multi-GPU split sit above it)
            const uint64_t space = h->prepart_space ? h->prepart_space : ((uint64_t)1 << h->b1);
            const uint64_t digit = (hs >> h->b2) & (space - 1);
            jp.sentinel_part = 0xffffffffu;
            if (!h->parent_digits.empty()) {
                for (uint32_t i = 0; i < h->d1; ++i)
                    if (h->parent_digits[i] == digit) jp.sentinel_part = (uint32_t)(i * h->d2 + (hs & (h->d2 - 1)));
            } else if (digit >= h->prepart_first && digit < (uint64_t)h->prepart_first + h->d1) {
                jp.sentinel_part = (uint32_t)((digit - h->prepart_first) * h->d2 + (hs & (h->d2 - 1)));
            }
        }
    }
    jp.matches = h->d_matches;
    jp.cta_times = h->d_cta_times;
    if (h->l2join) {
        PtParams q{};
        q.build = part_build;
        q.probe = part_probe;
        q.bounds_build = h->d_bounds2[0];
        q.bounds_probe = h->d_bounds2[1];
        q.npart = (uint32_t)h->nparts;
        q.max_keys = h->join_max_keys;
        q.table_mul = jp.table_mul;
        q.sentinel_part = jp.sentinel_part;
        q.table = h->d_pt;
        q.region_buckets = h->pt_region_buckets;
        q.region_shift32 = 32 - (uint32_t)ilog2_ceil(h->pt_region_buckets);
        q.flags = h->d_scalars + kGtFlags;
        q.matches = h->d_matches;
        const uint32_t grid = (uint32_t)h->sm_count * 8;
        {
            KernelScope ks(h, "pt_clear");
            PHJ_CUDA(cudaMemsetAsync(h->d_scalars + kGtFlags, 0, 4, h->stream));
            gt_clear<<<grid, 256, 0, h->stream>>>(h->d_pt, (uint64_t)h->nparts * h->pt_region_buckets * 4);
        }
        {
            KernelScope ks(h, "pt_build");
            pt_build<256><<<grid, 256, 0, h->stream>>>(q);
        }
        PHJ_CUDA(cudaEventRecord(h->ev[2], h->stream));
        {
            // exactly the CTAs that are resident at once (256 threads x 40 registers: six per SM): the tiles are
            // dealt round-robin, so a second wave would walk the whole probe side again behind the first
            KernelScope ks(h, "pt_probe");
            pt_probe<256><<<(uint32_t)h->sm_count * 6, 256, 0, h->stream>>>(q);
        }
        KernelScope ks(h, "count_oversize");
        count_oversize<<<(uint32_t)((h->nparts + 255) / 256), 256, 0, h->stream>>>(
            h->d_bounds2[0], (uint32_t)h->nparts, h->join_max_keys, d_oversize);
    } else {
        KernelScope ks(h, "join_partitions");
        jp.oversize = d_oversize;  // the join kernel also counts the partitions it has to leave to the global table
        PHJ_CUDA(launch_join(h, jp, h->join_grid, (size_t)h->join_slots * 8));
    }
    PHJ_CUDA(cudaEventRecord(h->ev[3], h->stream));
    // {matches, oversize | truncated << 32} in one copy; the per-CTA phase times (diagnostics) follow it
    PHJ_CUDA(cudaMemcpyAsync(h->h_out, h->d_matches, 16, cudaMemcpyDeviceToHost, h->stream));
    PHJ_CUDA(cudaEventRecord(h->ev[4], h->stream));
    PHJ_CUDA(cudaMemcpyAsync(h->h_cta_times, h->d_cta_times, (size_t)h->join_grid * 16,
                             cudaMemcpyDeviceToHost, h->stream));
    PHJ_CUDA(cudaStreamSynchronize(h->stream));
    PHJ_CUDA(cudaGetLastError());

    const uint32_t oversize = (uint32_t)(h->h_out[1] & 0xffffffffu);
    if (h->h_out[1] >> 32)  // plan_pass2's truncation flag shares the word
        return fail(PHJ_ERR_INVALID, "internal: the pass-2 plan needed more than %u segments", h->max_segs2);
    float extra_ms = 0;
    if (oversize && h->prepart && !h->parent_digits.empty())
        return fail(PHJ_ERR_INVALID, "%u build partitions exceed the shared-memory table; with explicitly numbered "
                                     "parents raise `partitions` instead", oversize);
    if (oversize) {
        // Build partitions too large for shared memory: join them through the global table.
        int rc = ensure_gt_for_fallback(h);
        if (rc != PHJ_OK) return rc;
        PHJ_CUDA(cudaEventRecord(h->ev[5], h->stream));
        if ((rc = run_gt(h, true, part_build, part_probe)) != PHJ_OK) return rc;
        PHJ_CUDA(cudaMemcpyAsync(h->h_out, h->d_matches, 8, cudaMemcpyDeviceToHost, h->stream));
        PHJ_CUDA(cudaEventRecord(h->ev[4], h->stream));
        PHJ_CUDA(cudaStreamSynchronize(h->stream));
        PHJ_CUDA(cudaGetLastError());
        extra_ms = ev_ms(h->ev[5], h->ev[4]);
    }
    out->matches = h->h_out[0];
    out->partition_ns = (uint64_t)(ev_ms(h->ev[0], h->ev[1]) * 1e6);
    out->join_ns = (uint64_t)((ev_ms(h->ev[1], h->ev[3]) + extra_ms) * 1e6);
    out->total_ns = (uint64_t)((ev_ms(h->ev[0], oversize ? h->ev[3] : h->ev[4]) + extra_ms) * 1e6);
    // The reference reports build/probe of the worker with the largest build+probe
    // (src/RadixCluster/HashJoin.hpp:67-87); a CTA is the worker here.
    uint64_t best_b = 0, best_p = 0;
    for (uint32_t i = 0; i < h->join_grid; ++i) {
        uint64_t b = h->h_cta_times[2 * i], p = h->h_cta_times[2 * i + 1];
        if (b + p > best_b + best_p) {
            best_b = b;
            best_p = p;
        }
    }
    out->build_ns = best_b;
    out->probe_ns = best_p;
    if (h->l2join) {  // the table kernels do not time themselves: whole-kernel device times instead
        out->build_ns = (uint64_t)(ev_ms(h->ev[1], h->ev[2]) * 1e6);
        out->probe_ns = (uint64_t)(ev_ms(h->ev[2], h->ev[3]) * 1e6);
    }
    out->passes = (two ? 2 : 1) - (h->prepart ? 1 : 0);
    out->partitions = h->P;
    out->fallback_partitions = oversize;
    out->d2h_bytes = 8 + 4 + (uint64_t)h->join_grid * 16 + (oversize ? 8 : 0);
    // one histogram read + per pass (read + write) + one join read, 16 B each
    out->hbm_bytes_alg = 16ull * (2 + 2 * out->passes) * (h->n[0] + h->n[1]);
    h->joined_radix = true;
    return PHJ_OK;
}

int validate_config(const phj_config* c) {
    if (!c) return fail(PHJ_ERR_INVALID, "config is null");
    if (c->algo != PHJ_ALGO_NO_PARTITIONING && c->algo != PHJ_ALGO_RADIX_PARTITIONING &&
        c->algo != PHJ_ALGO_SHARD_SPLIT)
        return fail(PHJ_ERR_INVALID, "Unrecognized join algorithm: %d.", c->algo);
    if (c->algo == PHJ_ALGO_SHARD_SPLIT) {
        const uint64_t w = c->partitions;
        // a digit count that is not a power of two (ranks x local partitions for 3, 5, 6, 7 ... ranks) is taken
        // as hash % partitions, which has no bit field to shift
        if (w == 0 || w > kMaxSplitDigits || ((w & (w - 1)) && c->shard_shift != 0) || c->shard_shift > 56 ||
            c->radix_bits[0] || c->radix_bits[1])
            return fail(PHJ_ERR_INVALID, "shard split needs partitions = ranks x local pass-1 digits (<= 256; a power "
                                         "of two unless shard_shift is 0), shard_shift <= 56 and no radix_bits");
        if (c->split_chunks > (uint32_t)kMaxSplitChunks)
            return fail(PHJ_ERR_INVALID, "split_chunks must be <= %d", kMaxSplitChunks);
    }
    if (c->num_gpus < 0) return fail(PHJ_ERR_INVALID, "num_gpus must be >= 0");
    if (c->upload_chunks > (uint32_t)kMaxUploadChunks)
        return fail(PHJ_ERR_INVALID, "upload_chunks must be <= %d", kMaxUploadChunks);
    if ((c->flags & PHJ_FLAG_CHAINED_TABLE) && c->algo != PHJ_ALGO_NO_PARTITIONING)
        return fail(PHJ_ERR_INVALID, "the bucket-chained table serves the no-partitioning join only");
    if (c->hash < PHJ_HASH_XXH3 || c->hash > PHJ_HASH_CITY)
        return fail(PHJ_ERR_INVALID, "Unrecognized hash function: %d.", c->hash);
    if (c->algo == PHJ_ALGO_NO_PARTITIONING && (c->partitions != 0 || c->radix_bits[0] || c->radix_bits[1]))
        // src/Arguments.hpp:12-17
        return fail(PHJ_ERR_INVALID,
                    "number of partitions can be specified only for RadixParitioning.");
    return PHJ_OK;
}

}  // namespace

// =================================================================================================
// C ABI
// =================================================================================================
extern "C" {

const char* phj_last_error(void) { return g_error.c_str(); }
uint32_t phj_abi_version(void) { return PHJ_ABI_VERSION; }

int phj_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

int phj_get_device_info(int32_t device, phj_device_info* out) {
    if (!out) return fail(PHJ_ERR_INVALID, "out is null");
    cudaDeviceProp prop;
    PHJ_CUDA(cudaGetDeviceProperties(&prop, device));
    memset(out, 0, sizeof(*out));
    snprintf(out->name, sizeof(out->name), "%.127s", prop.name);
    out->sm_count = prop.multiProcessorCount;
    out->cc_major = prop.major;
    out->cc_minor = prop.minor;
    out->global_mem_bytes = prop.totalGlobalMem;
    out->l2_bytes = (uint64_t)prop.l2CacheSize;
    out->smem_per_block_optin = prop.sharedMemPerBlockOptin;
    int v = 0;
    cudaDeviceGetAttribute(&v, cudaDevAttrClockRate, device);
    out->sm_clock_khz = v;
    cudaDeviceGetAttribute(&v, cudaDevAttrMemoryClockRate, device);
    out->mem_clock_khz = v;
    cudaDeviceGetAttribute(&v, cudaDevAttrGlobalMemoryBusWidth, device);
    out->mem_bus_bits = v;
    return PHJ_OK;
}

int phj_create(const phj_config* config, phj_handle** out) {
    if (!out) return fail(PHJ_ERR_INVALID, "out is null");
    *out = nullptr;
    int rc = validate_config(config);
    if (rc != PHJ_OK) return rc;
    if (config->num_gpus > 1) {  // the sharded join over the GPUs of this node, driven from this process
        phj_handle* g = new phj_handle;
        g->cfg = *config;
        g->device = config->device;
        if ((rc = group_create(config, &g->group)) != PHJ_OK) {
            delete g;
            return rc;
        }
        *out = g;
        return PHJ_OK;
    }
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        return fail(PHJ_ERR_CUDA, "no CUDA device available (%s); this engine has no CPU fallback",
                    e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
    }
    if (config->device < 0 || config->device >= ndev)
        return fail(PHJ_ERR_INVALID, "device %d out of range [0, %d)", config->device, ndev);
    PHJ_CUDA(cudaSetDevice(config->device));
    cudaDeviceProp prop;
    PHJ_CUDA(cudaGetDeviceProperties(&prop, config->device));
    if (prop.major < 10)
        return fail(PHJ_ERR_CUDA, "device %d is sm_%d%d; this library is built for sm_100a only",
                    config->device, prop.major, prop.minor);
    phj_handle* h = new phj_handle;
    h->cfg = *config;
    h->device = config->device;
    h->sm_count = prop.multiProcessorCount;
    h->smem_optin = prop.sharedMemPerBlockOptin;
    // PHJ_KERNEL_TIMES=1: CUDA events around every kernel; any other non-empty value: only around kernels whose
    // name contains it (every pair of events costs the stream about a microsecond)
    const char* kt = getenv("PHJ_KERNEL_TIMES");
    h->time_kernels = kt && kt[0] && kt[0] != '0';
    if (h->time_kernels && strcmp(kt, "1")) h->ktime_filter = kt;
    if (const char* rk = getenv("PHJ_RANK")) {
        h->use_match = !strcmp(rk, "match");
        h->use_lanes = strcmp(rk, "match") && strcmp(rk, "stable");
    }
    const char* ds = getenv("PHJ_DEBUG_SYNC");
    h->debug_sync = ds && ds[0] == '1';
    auto cleanup = [&](int code) {
        phj_destroy(h);
        return code;
    };
    int prio_lo = 0, prio_hi = 0;
    cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);
    // the multi-GPU split runs beside local joins: its CTAs go first whenever an SM slot frees up
    if (cudaStreamCreateWithPriority(&h->stream, cudaStreamNonBlocking,
                                     config->algo == PHJ_ALGO_SHARD_SPLIT ? prio_hi : prio_lo) != cudaSuccess)
        return cleanup(fail(PHJ_ERR_CUDA, "cudaStreamCreate failed"));
    for (auto& ev : h->ev)
        if (cudaEventCreate(&ev) != cudaSuccess) return cleanup(fail(PHJ_ERR_CUDA, "cudaEventCreate failed"));
    for (auto& k : h->ktimes) {
        cudaEventCreate(&k.begin);
        cudaEventCreate(&k.end);
    }
    if (cudaMalloc(&h->d_scalars, kNumScalars * 4) != cudaSuccess ||
        cudaMalloc(&h->d_matches, 32) != cudaSuccess ||
        cudaMallocHost(&h->h_out, 64) != cudaSuccess)
        return cleanup(fail(PHJ_ERR_NOMEM, "allocation of engine scalars failed"));
    cudaMemset(h->d_scalars, 0, kNumScalars * 4);
    *out = h;
    return PHJ_OK;
}

void phj_destroy(phj_handle* h) {
    if (!h) return;
    if (h->group) {
        group_destroy(h->group);
        delete h;
        return;
    }
    cudaSetDevice(h->device);
    if (h->stream_child) phj_destroy(h->stream_child);
    if (h->upload_stream) {
        cudaStreamSynchronize(h->upload_stream);
        cudaStreamDestroy(h->upload_stream);
    }
    for (auto& e : h->upload_ev)
        if (e) cudaEventDestroy(e);
    if (h->stream) cudaStreamSynchronize(h->stream);
    for (int rel = 0; rel < 2; ++rel) {
        if (h->owns_in[rel] && h->d_in[rel]) cudaFree(h->d_in[rel]);
        if (h->d_buf_a[rel]) cudaFree(h->d_buf_a[rel]);
        if (h->d_buf_b[rel]) cudaFree(h->d_buf_b[rel]);
        if (h->d_bounds1[rel]) cudaFree(h->d_bounds1[rel]);
        if (h->d_bounds2[rel]) cudaFree(h->d_bounds2[rel]);
        if (h->d_parents2[rel]) cudaFree(h->d_parents2[rel]);
    }
    void* ptrs[] = {h->d_segs1, h->d_segs2, h->d_scalars, h->d_counts, h->d_cursors,
                    h->d_scan_state, h->d_matches, h->d_cta_times, h->d_gt, h->d_hist12, h->d_pt, h->d_outd[0], h->d_outd[1], h->d_pre_bounds, h->d_shard_starts, h->d_joined, h->d_cta_rows, h->d_ct_heads, h->d_ct_buckets};
    for (void* p : ptrs)
        if (p) cudaFree(p);
    if (h->h_out) cudaFreeHost(h->h_out);
    if (h->h_cta_times) cudaFreeHost(h->h_cta_times);
    if (h->h_shard_starts) cudaFreeHost(h->h_shard_starts);
    for (auto& ev : h->ev)
        if (ev) cudaEventDestroy(ev);
    for (auto& k : h->ktimes) {
        if (k.begin) cudaEventDestroy(k.begin);
        if (k.end) cudaEventDestroy(k.end);
    }
    if (h->stream) cudaStreamDestroy(h->stream);
    delete h;
}

static int set_relations(phj_handle* h, const void* build, size_t n_build, const void* probe,
                         size_t n_probe, bool device_resident, uint64_t* h2d_ns, bool prepartitioned = false,
                         uint32_t parents = 0, bool defer_copy = false) {
    if (!h) return fail(PHJ_ERR_INVALID, "handle is null");
    if ((n_build && !build) || (n_probe && !probe))
        return fail(PHJ_ERR_INVALID, "relation pointer is null but its size is not zero");
    if (((uintptr_t)build | (uintptr_t)probe) & 15)
        return fail(PHJ_ERR_INVALID, "relations must be 16-byte aligned (alignas(16) Tuple)");
    if (h->group) {
        if (prepartitioned) return fail(PHJ_ERR_STATE, "num_gpus > 1: pre-partitioned relations are not supported");
        int rc = group_set_relations(h->group, build, n_build, probe, n_probe, device_resident);
        if (rc == PHJ_OK) {
            h->n[0] = n_build;
            h->n[1] = n_probe;
            h->have_data = true;
        }
        return rc;
    }
    PHJ_CUDA(cudaSetDevice(h->device));
    const void* src[2] = {build, probe};
    const size_t nn[2] = {n_build, n_probe};
    // A pre-partitioned bind (shrink_ok) keeps a plan that has room; everything else needs the pass-1
    // segment table to cover exactly these sizes.
    const bool replan = !h->have_data || nn[0] > h->plan_n[0] || nn[1] > h->plan_n[1] ||
                        parents != h->prepart_parents ||
                        (!prepartitioned && (nn[0] != h->segs_n[0] || nn[1] != h->segs_n[1]));
    h->prepart_parents = parents;
    const bool resized = nn[0] != h->n[0] || nn[1] != h->n[1];
    for (int rel = 0; rel < 2; ++rel) {
        if (device_resident) {
            if (h->owns_in[rel] && h->d_in[rel]) cudaFree(h->d_in[rel]);
            h->owns_in[rel] = false;
            h->cap_in[rel] = 0;
            h->d_in[rel] = reinterpret_cast<ulonglong2*>(const_cast<void*>(src[rel]));
        } else {
            if (!h->owns_in[rel]) {
                h->d_in[rel] = nullptr;
                h->cap_in[rel] = 0;
            }
            size_t want = std::max<size_t>(std::max<size_t>(nn[rel], 1),
                                           rel == 0 ? h->cfg.reserve_build : h->cfg.reserve_probe);
            if (nn[rel] > h->cap_in[rel] || !h->d_in[rel]) {
                if (h->d_in[rel]) cudaFree(h->d_in[rel]);
                h->d_in[rel] = nullptr;
                PHJ_CUDA(cudaMalloc(&h->d_in[rel], want * 16));
                h->cap_in[rel] = want;
            }
            h->owns_in[rel] = true;
        }
        h->n[rel] = nn[rel];
    }
    if (replan) {
        h->have_data = false;
        int rc = build_plan(h);
        if (rc != PHJ_OK) return rc;
    } else if (resized && h->cfg.algo != PHJ_ALGO_NO_PARTITIONING) {
        int rc = upload_tail_bounds(h);
        if (rc != PHJ_OK) return rc;
    }
    if (!device_resident && !defer_copy) {  // defer_copy: the streamed phj_join_host uploads by itself
        PHJ_CUDA(cudaEventRecord(h->ev[5], h->stream));
        for (int rel = 0; rel < 2; ++rel)
            if (nn[rel])
                PHJ_CUDA(cudaMemcpyAsync(h->d_in[rel], src[rel], nn[rel] * 16, cudaMemcpyHostToDevice,
                                         h->stream));
        PHJ_CUDA(cudaEventRecord(h->ev[0], h->stream));
        PHJ_CUDA(cudaStreamSynchronize(h->stream));
        if (h2d_ns) *h2d_ns = (uint64_t)(ev_ms(h->ev[5], h->ev[0]) * 1e6);
    }
    h->have_data = true;
    h->joined_radix = false;
    h->prepart = false;
    if (!prepartitioned && h->cfg.algo != PHJ_ALGO_NO_PARTITIONING &&
        !(h->cfg.algo == PHJ_ALGO_SHARD_SPLIT && (h->cfg.flags & PHJ_FLAG_SPLIT_REMOTE_ONLY))) {
        int rc = ensure_buffers(h, true, h->b2 > 0);
        if (rc != PHJ_OK) return rc;
    }
    return PHJ_OK;
}

int phj_upload(phj_handle* h, const phj_tuple* build, size_t n_build, const phj_tuple* probe,
               size_t n_probe) {
    return set_relations(h, build, n_build, probe, n_probe, false, nullptr);
}

int phj_bind_device(phj_handle* h, const void* d_build, size_t n_build, const void* d_probe,
                    size_t n_probe) {
    return set_relations(h, d_build, n_build, d_probe, n_probe, true, nullptr);
}

int phj_bind_device_partitioned(phj_handle* h, const void* d_build, size_t n_build, const void* d_probe,
                                size_t n_probe, const uint64_t* bounds_build, const uint64_t* bounds_probe,
                                uint32_t nparents, uint32_t first_parent, uint32_t parent_space) {
    if (!h) return fail(PHJ_ERR_INVALID, "handle is null");
    if (parent_space & (parent_space - 1)) return fail(PHJ_ERR_INVALID, "parent_space must be a power of two (or 0)");
    if (h->cfg.algo != PHJ_ALGO_RADIX_PARTITIONING)
        return fail(PHJ_ERR_STATE, "pre-partitioned relations need a radix-partitioning handle");
    if (!bounds_build || !bounds_probe) return fail(PHJ_ERR_INVALID, "partition boundaries are null");
    const uint64_t P = h->cfg.partitions;
    if (P == 0 || (P & (P - 1)))
        return fail(PHJ_ERR_INVALID, "pre-partitioned relations need an explicit power-of-two `partitions`");
    if (nparents == 0 || nparents > (uint32_t)kMaxSplitDigits)
        return fail(PHJ_ERR_INVALID, "nparents must be in [1, %d]", kMaxSplitDigits);
    const uint64_t* hb[2] = {bounds_build, bounds_probe};
    const size_t nn[2] = {n_build, n_probe};
    for (int rel = 0; rel < 2; ++rel) {
        if (hb[rel][0] != 0 || hb[rel][nparents] != nn[rel])
            return fail(PHJ_ERR_INVALID, "boundaries must start at 0 and end at the relation size");
        for (uint32_t d = 0; d < nparents; ++d)
            if (hb[rel][d] > hb[rel][d + 1]) return fail(PHJ_ERR_INVALID, "boundaries must not decrease");
    }
    int rc = set_relations(h, d_build, n_build, d_probe, n_probe, true, nullptr, true, nparents);
    if (rc != PHJ_OK) return rc;
    if ((rc = ensure_buffers(h, false, h->b2 > 0)) != PHJ_OK) return rc;
    h->prepart_first = first_parent;
    h->prepart_space = parent_space;
    h->parent_digits.clear();
    if ((rc = dev_reserve(&h->d_pre_bounds, &h->cap_pre_bounds, 2 * ((size_t)h->d1 + 1))) != PHJ_OK) return rc;
    for (int rel = 0; rel < 2; ++rel)
        PHJ_CUDA(cudaMemcpyAsync(h->d_pre_bounds + rel * (h->d1 + 1), hb[rel], (h->d1 + 1) * 8,
                                 cudaMemcpyHostToDevice, h->stream));
    PHJ_CUDA(cudaStreamSynchronize(h->stream));
    h->prepart = true;
    return PHJ_OK;
}

int phj_set_parent_digits(phj_handle* h, const uint32_t* digits, uint32_t n) {
    if (!h) return fail(PHJ_ERR_INVALID, "handle is null");
    if (!h->prepart) return fail(PHJ_ERR_STATE, "phj_set_parent_digits follows phj_bind_device_partitioned");
    if (n != h->d1 || !digits) return fail(PHJ_ERR_INVALID, "expected the digit of each of the %u bound parents", h->d1);
    h->parent_digits.assign(digits, digits + n);
    return PHJ_OK;
}

int phj_memcpy_d2d(int32_t device, void* d_dst, const void* d_src, size_t bytes) {
    if (bytes && (!d_dst || !d_src)) return fail(PHJ_ERR_INVALID, "null argument");
    PHJ_CUDA(cudaSetDevice(device));
    // A device-to-device cudaMemcpy returns before the copy has run, and the engine's streams are
    // non-blocking (not ordered behind the legacy stream): wait here, so that a join launched next
    // sees the data.
    if (bytes) {
        PHJ_CUDA(cudaMemcpyAsync(d_dst, d_src, bytes, cudaMemcpyDeviceToDevice, cudaStreamPerThread));
        PHJ_CUDA(cudaStreamSynchronize(cudaStreamPerThread));
    }
    return PHJ_OK;
}

int phj_join(phj_handle* h, phj_result* out) {
    if (!h || !out) return fail(PHJ_ERR_INVALID, "handle or result is null");
    if (!h->have_data) return fail(PHJ_ERR_STATE, "phj_join called before phj_upload / phj_bind_device");
    if (h->group) return group_join(h->group, out);
    PHJ_CUDA(cudaSetDevice(h->device));
    memset(out, 0, sizeof(*out));
    h->launches = 0;
    h->n_ktimes = 0;
    int rc = h->cfg.algo == PHJ_ALGO_NO_PARTITIONING ? join_no_partitioning(h, out) : join_radix(h, out);
    out->kernel_launches = h->launches;
    return rc;
}

int phj_join_materialize(phj_handle* h, phj_result* out) {
    if (!h || !out) return fail(PHJ_ERR_INVALID, "handle or result is null");
    if (h->cfg.algo != PHJ_ALGO_RADIX_PARTITIONING)
        return fail(PHJ_ERR_STATE, "the joined table is produced by the radix-partitioning joiner");
    if (h->group) return fail(PHJ_ERR_STATE, "num_gpus > 1 counts only: the joined table needs a single-GPU handle");
    int rc = phj_join(h, out);  // partitions both relations (and counts, as the reference does)
    if (rc != PHJ_OK) return rc;
    const bool two = h->b2 > 0;
    constexpr int kTpb = 512;
    MatParams mp{};
    mp.build = two ? h->d_buf_b[0] : h->prepart ? h->d_in[0] : h->d_buf_a[0];
    mp.probe = two ? h->d_buf_b[1] : h->prepart ? h->d_in[1] : h->d_buf_a[1];
    mp.bounds_build = h->d_bounds2[0];
    mp.bounds_probe = h->d_bounds2[1];
    mp.n_probe = h->n[1];
    mp.npart = (uint32_t)h->nparts;
    // table chunk: 4096 build tuples (64 KB) + 8192 index slots (32 KB): two CTAs per SM
    mp.cap_tuples = 4096;
    const uint32_t slots = 8192;
    mp.slot_mask = slots - 1;
    mp.slot_shift = 64 - ilog2_ceil(slots);
    mp.table_mul = (h->cfg.table_seed * 0x9E3779B97F4A7C15ULL) | 1ULL;
    if (h->cfg.table_seed == 0) mp.table_mul = 0xBF58476D1CE4E5B9ULL;
    const uint32_t grid = std::min<uint32_t>((uint32_t)h->sm_count * 2 * 4, 8192);
    mp.slice_len = h->n[1] / grid;
    mp.slice_rem = h->n[1] % grid;
    mp.cursor = h->d_matches + 2;
    if (!h->d_cta_rows) PHJ_CUDA(cudaMalloc(&h->d_cta_rows, 8192 * sizeof(unsigned long long)));
    mp.cta_rows = h->d_cta_rows;
    const size_t smem = (size_t)mp.cap_tuples * 16 + (size_t)slots * 4;
    auto count_kern = join_materialize<kTpb, false>;
    auto write_kern = join_materialize<kTpb, true>;
    PHJ_CUDA(cudaFuncSetAttribute(count_kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    PHJ_CUDA(cudaFuncSetAttribute(write_kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    PHJ_CUDA(cudaEventRecord(h->ev[5], h->stream));
    PHJ_CUDA(cudaMemsetAsync(h->d_matches + 2, 0, 8, h->stream));
    {
        KernelScope ks(h, "join_materialize[count]");
        count_kern<<<grid, kTpb, smem, h->stream>>>(mp);
    }
    PHJ_CUDA(cudaMemcpyAsync(h->h_out + 2, h->d_matches + 2, 8, cudaMemcpyDeviceToHost, h->stream));
    PHJ_CUDA(cudaStreamSynchronize(h->stream));
    PHJ_CUDA(cudaGetLastError());
    const uint64_t rows = h->h_out[2];
    if (rows > h->cap_joined) {
        if (h->d_joined) cudaFree(h->d_joined);
        h->d_joined = nullptr;
        h->cap_joined = 0;
        PHJ_CUDA(cudaMalloc(&h->d_joined, rows * 24));
        h->cap_joined = rows;
    }
    mp.out = h->d_joined;
    mp.out_cap = h->cap_joined;
    if (rows) {
        KernelScope ks(h, "join_materialize[write]", 2);
        scan_cta_rows<<<1, 1024, 0, h->stream>>>(h->d_cta_rows, grid);
        write_kern<<<grid, kTpb, smem, h->stream>>>(mp);
    }
    PHJ_CUDA(cudaEventRecord(h->ev[4], h->stream));
    PHJ_CUDA(cudaStreamSynchronize(h->stream));
    PHJ_CUDA(cudaGetLastError());
    h->n_joined = rows;
    out->joined_tuples = rows;
    out->materialize_ns = (uint64_t)(ev_ms(h->ev[5], h->ev[4]) * 1e6);
    out->total_ns += out->materialize_ns;
    out->kernel_launches = h->launches;
    out->d2h_bytes += 16;
    // probe read twice + build read + rows written
    out->hbm_bytes_alg += 2 * 16ull * h->n[1] + 2 * 16ull * h->n[0] + 24ull * rows;
    return PHJ_OK;
}

int phj_read_joined(phj_handle* h, phj_joined_tuple* out, uint64_t first, uint64_t count) {
    if (!h) return fail(PHJ_ERR_INVALID, "handle is null");
    if (first + count > h->n_joined) return fail(PHJ_ERR_INVALID, "rows [%llu, %llu) exceed the joined table (%llu rows)",
                                                 (unsigned long long)first, (unsigned long long)(first + count),
                                                 (unsigned long long)h->n_joined);
    if (count && !out) return fail(PHJ_ERR_INVALID, "out is null");
    PHJ_CUDA(cudaSetDevice(h->device));
    if (count) PHJ_CUDA(cudaMemcpy(out, h->d_joined + first * 3, count * 24, cudaMemcpyDeviceToHost));
    return PHJ_OK;
}

int phj_device_joined(phj_handle* h, const void** d_joined, uint64_t* rows) {
    if (!h) return fail(PHJ_ERR_INVALID, "handle is null");
    if (d_joined) *d_joined = h->d_joined;
    if (rows) *rows = h->n_joined;
    return PHJ_OK;
}

// Streamed host join: the build relation and then the probe relation in `chunks` row chunks go up
// on their own stream (issued by a helper thread, so that pageable memory, whose "async" copies
// block the issuing thread, overlaps as well); as soon as a chunk has landed the child handle joins
// it against the whole build relation. Counts are additive over a partition of the probe relation,
// so the sum is the count of the whole join (SURVEY.md 8d uses the same identity for config 5). The
// build side is re-partitioned for every chunk: redundant work on SMs that would otherwise idle
// behind PCIe. Only the last chunk's join is not hidden behind the upload.
static int join_host_streamed(phj_handle* h, const phj_tuple* build, size_t n_build, const phj_tuple* probe,
                              size_t n_probe, uint32_t chunks, phj_result* out) {
    int rc = set_relations(h, build, n_build, probe, n_probe, false, nullptr, false, 0, /*defer_copy=*/true);
    if (rc != PHJ_OK) return rc;
    h->have_data = false;  // until the uploads below have landed
    const size_t chunk_len = (n_probe + chunks - 1) / chunks;
    if (!h->stream_child) {
        phj_config cc = h->cfg;
        cc.upload_chunks = 1;
        cc.reserve_build = std::max<uint64_t>(cc.reserve_build, n_build);
        cc.reserve_probe = chunk_len;
        if ((rc = phj_create(&cc, &h->stream_child)) != PHJ_OK) return rc;
    }
    if (!h->upload_stream) {
        PHJ_CUDA(cudaStreamCreateWithFlags(&h->upload_stream, cudaStreamNonBlocking));
        for (auto& e : h->upload_ev) PHJ_CUDA(cudaEventCreate(&e));
    }
    phj_handle* child = h->stream_child;

    struct Uploader {
        std::atomic<int> published{0};  // chunks whose "landed" event has been recorded
        std::atomic<bool> failed{false};
        std::string error;
        std::thread thread;
        ~Uploader() {
            if (thread.joinable()) thread.join();
        }
    } up;
    auto uploader = [&]() {
        auto check = [&](cudaError_t e, const char* what) {
            if (e == cudaSuccess) return true;
            up.error = std::string(what) + " failed: " + cudaGetErrorString(e);
            up.failed.store(true);
            return false;
        };
        if (!check(cudaSetDevice(h->device), "cudaSetDevice")) return;
        if (!check(cudaEventRecord(h->upload_ev[0], h->upload_stream), "cudaEventRecord")) return;
        if (n_build && !check(cudaMemcpyAsync(h->d_in[0], build, n_build * 16, cudaMemcpyHostToDevice,
                                              h->upload_stream), "upload of the build relation")) return;
        if (!check(cudaEventRecord(h->upload_ev[1], h->upload_stream), "cudaEventRecord")) return;
        for (uint32_t c = 0; c < chunks; ++c) {
            const size_t first = std::min(n_probe, (size_t)c * chunk_len);
            const size_t len = std::min(n_probe - first, chunk_len);
            if (len && !check(cudaMemcpyAsync(h->d_in[1] + first, probe + first, len * 16, cudaMemcpyHostToDevice,
                                              h->upload_stream), "upload of a probe chunk")) return;
            if (!check(cudaEventRecord(h->upload_ev[2 + c], h->upload_stream), "cudaEventRecord")) return;
            up.published.store((int)c + 1, std::memory_order_release);
        }
    };
    try {
        up.thread = std::thread(uploader);
    } catch (const std::exception&) {
        uploader();  // no helper thread to be had: issue the uploads from here (pinned memory still overlaps)
    }

    memset(out, 0, sizeof(*out));
    auto join_chunks = [&]() -> int {
        for (uint32_t c = 0; c < chunks; ++c) {
            const size_t first = std::min(n_probe, (size_t)c * chunk_len);
            const size_t len = std::min(n_probe - first, chunk_len);
            // plan for this chunk while it is still on its way
            int rc2 = phj_bind_device(child, h->d_in[0], n_build, h->d_in[1] + first, len);
            if (rc2 != PHJ_OK) return rc2;
            while (up.published.load(std::memory_order_acquire) <= (int)c && !up.failed.load()) std::this_thread::yield();
            if (up.failed.load()) return PHJ_OK;  // reported below, with the uploader's message
            PHJ_CUDA(cudaStreamWaitEvent(child->stream, h->upload_ev[2 + c], 0));
            phj_result r;
            if ((rc2 = phj_join(child, &r)) != PHJ_OK) return rc2;
            out->matches += r.matches;
            out->partition_ns += r.partition_ns;
            out->build_ns += r.build_ns;
            out->probe_ns += r.probe_ns;
            out->join_ns += r.join_ns;
            out->total_ns += r.total_ns;
            out->hbm_bytes_alg += r.hbm_bytes_alg;
            out->kernel_launches += r.kernel_launches;
            out->d2h_bytes += r.d2h_bytes;
            out->fallback_partitions = std::max(out->fallback_partitions, r.fallback_partitions);
            out->passes = r.passes;
            out->partitions = r.partitions;
        }
        return PHJ_OK;
    };
    rc = join_chunks();
    // Every exit goes through here: the uploader has issued all it will issue, and its copies -- which read the
    // CALLER's buffers -- have drained before the caller gets control back, error or not.
    const std::string join_error = g_error;
    if (up.thread.joinable()) up.thread.join();
    cudaStreamSynchronize(h->upload_stream);
    cudaStreamSynchronize(child->stream);
    if (rc != PHJ_OK) {
        g_error = join_error;
        return rc;
    }
    if (up.failed.load()) return fail(PHJ_ERR_CUDA, "%s", up.error.c_str());
    PHJ_CUDA(cudaGetLastError());
    PHJ_CUDA(cudaEventRecord(h->ev[4], child->stream));
    PHJ_CUDA(cudaEventSynchronize(h->ev[4]));
    out->h2d_ns = (uint64_t)(ev_ms(h->upload_ev[0], h->upload_ev[1 + chunks]) * 1e6);
    out->e2e_ns = (uint64_t)(ev_ms(h->upload_ev[0], h->ev[4]) * 1e6);
    out->h2d_bytes = 16ull * (n_build + n_probe);
    out->upload_chunks = chunks;
    h->have_data = true;  // both relations are resident: phj_join may follow
    return PHJ_OK;
}

int phj_join_host(phj_handle* h, const phj_tuple* build, size_t n_build, const phj_tuple* probe,
                  size_t n_probe, phj_result* out) {
    if (!h || !out) return fail(PHJ_ERR_INVALID, "handle or result is null");
    if (h->group) {  // every GPU uploads its row shard (concurrently), then the sharded join
        const auto t0 = std::chrono::steady_clock::now();
        int rc = group_set_relations(h->group, build, n_build, probe, n_probe, false);
        if (rc != PHJ_OK) return rc;
        h->have_data = true;
        const auto t1 = std::chrono::steady_clock::now();
        if ((rc = group_join(h->group, out)) != PHJ_OK) return rc;
        out->h2d_ns = (uint64_t)std::chrono::duration_cast<std::chrono::nanoseconds>(t1 - t0).count();
        out->e2e_ns = (uint64_t)std::chrono::duration_cast<std::chrono::nanoseconds>(
                          std::chrono::steady_clock::now() - t0).count();
        out->h2d_bytes = 16ull * (n_build + n_probe);
        out->upload_chunks = 1;
        return PHJ_OK;
    }
    uint32_t chunks = h->cfg.upload_chunks;
    if (chunks == 0) {
        // automatic: chunks of ~256 MB once the probe relation is worth overlapping, but never so short
        // that redoing the build side per chunk (radix: partitioning it, ~14x faster per tuple than
        // PCIe delivers tuples; no-partitioning: building the table) takes longer than the chunk's upload
        const size_t by_bytes = (n_probe * 16 + kUploadChunkBytes - 1) / kUploadChunkBytes;
        const size_t min_len = std::max<size_t>(1, h->cfg.algo == PHJ_ALGO_NO_PARTITIONING ? n_build : n_build / 4);
        chunks = (uint32_t)std::min<size_t>(std::min<size_t>(kMaxUploadChunks, by_bytes), n_probe / min_len);
    }
    if (h->cfg.algo == PHJ_ALGO_SHARD_SPLIT || n_probe < chunks) chunks = 1;
    if (chunks > 1) return join_host_streamed(h, build, n_build, probe, n_probe, chunks, out);
    uint64_t h2d = 0;
    int rc = set_relations(h, build, n_build, probe, n_probe, false, &h2d);
    if (rc != PHJ_OK) return rc;
    rc = phj_join(h, out);
    if (rc == PHJ_OK) {
        out->h2d_ns = h2d;
        out->e2e_ns = h2d + out->total_ns;
        out->h2d_bytes = 16ull * (n_build + n_probe);
        out->upload_chunks = 1;
    }
    return rc;
}

int phj_read_partitions(phj_handle* h, int32_t which, phj_tuple* out, uint64_t* bounds) {
    if (!h) return fail(PHJ_ERR_INVALID, "handle is null");
    if (which < 0 || which > 1) return fail(PHJ_ERR_INVALID, "which must be 0 (build) or 1 (probe)");
    if (h->cfg.algo == PHJ_ALGO_NO_PARTITIONING || !h->joined_radix)
        return fail(PHJ_ERR_STATE, "no partitioned relations: run a radix-partitioning phj_join first");
    PHJ_CUDA(cudaSetDevice(h->device));
    const ulonglong2* src = h->b2 > 0 ? h->d_buf_b[which] : h->prepart ? h->d_in[which] : h->d_buf_a[which];
    if (out && h->n[which])
        PHJ_CUDA(cudaMemcpy(out, src, h->n[which] * 16, cudaMemcpyDeviceToHost));
    if (bounds) PHJ_CUDA(cudaMemcpy(bounds, h->d_bounds2[which], (h->P + 1) * 8, cudaMemcpyDeviceToHost));
    return PHJ_OK;
}

int phj_device_partitions(phj_handle* h, int32_t which, const void** d_data, const uint64_t** d_bounds,
                          size_t* n) {
    if (!h) return fail(PHJ_ERR_INVALID, "handle is null");
    if (which < 0 || which > 1) return fail(PHJ_ERR_INVALID, "which must be 0 (build) or 1 (probe)");
    if (h->cfg.algo == PHJ_ALGO_NO_PARTITIONING || !h->joined_radix)
        return fail(PHJ_ERR_STATE, "no partitioned relations: run a radix-partitioning phj_join first");
    if (d_data) *d_data = h->b2 > 0 ? h->d_buf_b[which] : h->prepart ? h->d_in[which] : h->d_buf_a[which];
    if (d_bounds) *d_bounds = h->d_bounds2[which];
    if (n) *n = h->n[which];
    return PHJ_OK;
}

// ---- fused multi-GPU shuffle -----------------------------------------------------------------
static void fill_pass1_params(phj_handle* h, PassParams& p1) {
    const HashParams hp = make_hash_params(h->cfg.hash, h->cfg.hash_seed);
    for (int rel = 0; rel < 2; ++rel) {
        p1.in[rel] = h->d_in[rel];
        p1.out[rel] = h->d_buf_a[rel];
        p1.bounds[rel] = h->d_bounds2[rel];
    }
    p1.segs = h->d_segs1;
    p1.nsegs = h->d_scalars + kNsegs1;
    p1.counts = h->d_counts;
    p1.cursors = h->d_cursors;
    p1.cursor_bias[0] = 0;
    p1.cursor_bias[1] = h->n[0];
    p1.bounds_stride = h->d1;
    p1.ndigits = h->d1;
    p1.hp = hp;
    p1.df = digit_fn(h, 1);
}

// Histogram + scan of the bound shard by split digit and the chunk starts of every digit, enqueued on
// the handle's stream: d_shard_starts[(rel * digits + d) * (chunks + 1) + c] = where chunk c's tuples
// of digit d start in the (virtual) digit-major split output of the relation; entry `chunks` is the
// end of the digit. Nothing here waits for the device.
static int shard_count_enqueue(phj_handle* h) {
    const uint32_t w = h->d1, K = h->nchunks;
    const size_t nstarts = (size_t)2 * w * (K + 1);
    if (!h->d_shard_starts) {
        const size_t cap = (size_t)2 * kMaxSplitDigits * (kMaxSplitChunks + 1);
        PHJ_CUDA(cudaMalloc(&h->d_shard_starts, cap * 8));
        PHJ_CUDA(cudaMallocHost(&h->h_shard_starts, cap * 8));
    }
    PassParams p1{};
    fill_pass1_params(h, p1);
    if (h->nsegs1 > 0) {
        KernelScope ks(h, "radix_histogram[split]");
        PHJ_CUDA(launch_pass(h, false, h->b1, p1, h->nsegs1));
    }
    if (h->nsegs1 > 0) run_scan(h, kNcounts1, (size_t)h->nsegs1 * h->d1);
    SplitStartsParams sp{};
    sp.cursors = h->d_cursors;
    sp.starts = h->d_shard_starts;
    for (int rel = 0; rel < 2; ++rel) {
        sp.cnt_base[rel] = h->cnt_base1_rel[rel];
        sp.nseg[rel] = h->nseg1_rel[rel];
        sp.n[rel] = h->n[rel];
        for (uint32_t c = 0; c <= K; ++c) sp.first_seg[rel][c] = h->chunk_first_seg[rel][c];
    }
    sp.bias[1] = h->n[0];
    sp.ndigits = w;
    sp.nchunks = K;
    {
        KernelScope ks(h, "split_starts");
        split_starts<<<(uint32_t)((nstarts + 255) / 256), 256, 0, h->stream>>>(sp);
    }
    return PHJ_OK;
}

int phj_shard_count(phj_handle* h, uint64_t* counts) {
    if (!h || !counts) return fail(PHJ_ERR_INVALID, "handle or counts is null");
    if (h->cfg.algo != PHJ_ALGO_SHARD_SPLIT) return fail(PHJ_ERR_STATE, "not a shard-split handle");
    if (!h->have_data) return fail(PHJ_ERR_STATE, "phj_shard_count called before phj_upload / phj_bind_device");
    PHJ_CUDA(cudaSetDevice(h->device));
    h->launches = 0;
    h->n_ktimes = 0;
    const uint32_t w = h->d1, K = h->nchunks;
    const size_t nstarts = (size_t)2 * w * (K + 1);
    PHJ_CUDA(cudaEventRecord(h->ev[0], h->stream));
    int rc = shard_count_enqueue(h);
    if (rc != PHJ_OK) return rc;
    PHJ_CUDA(cudaMemcpyAsync(h->h_shard_starts, h->d_shard_starts, nstarts * 8, cudaMemcpyDeviceToHost, h->stream));
    PHJ_CUDA(cudaStreamSynchronize(h->stream));
    PHJ_CUDA(cudaGetLastError());
    for (uint32_t c = 0; c < K; ++c)
        for (int rel = 0; rel < 2; ++rel)
            for (uint32_t d = 0; d < w; ++d) {
                const uint64_t* st = h->h_shard_starts + ((size_t)rel * w + d) * (K + 1);
                counts[((size_t)c * 2 + rel) * w + d] = st[c + 1] - st[c];
            }
    h->shard_counted = true;
    return PHJ_OK;
}

int phj_shard_scatter(phj_handle* h, uint32_t chunk, void* const* dst_build, const uint64_t* off_build,
                      void* const* dst_probe, const uint64_t* off_probe, phj_result* out) {
    if (!h || !out) return fail(PHJ_ERR_INVALID, "handle or result is null");
    if (h->cfg.algo != PHJ_ALGO_SHARD_SPLIT) return fail(PHJ_ERR_STATE, "not a shard-split handle");
    if (!h->shard_counted) return fail(PHJ_ERR_STATE, "phj_shard_scatter needs a preceding phj_shard_count");
    if (chunk >= h->nchunks) return fail(PHJ_ERR_INVALID, "chunk %u out of range [0, %u)", chunk, h->nchunks);
    PHJ_CUDA(cudaSetDevice(h->device));
    memset(out, 0, sizeof(*out));
    const uint32_t w = h->d1, K = h->nchunks;
    void* const* dst[2] = {dst_build, dst_probe};
    const uint64_t* off[2] = {off_build, off_probe};
    if (!dst_build || !dst_probe) {
        int rc = ensure_buffers(h, true, false);
        if (rc != PHJ_OK) return rc;
    }
    PassParams p1{};
    fill_pass1_params(h, p1);
    ulonglong2* host_ptrs[2][kMaxSplitDigits];
    for (int rel = 0; rel < 2; ++rel) {
        if (!dst[rel]) continue;  // this relation stays local (split into buf_a)
        if (!h->d_outd[rel]) PHJ_CUDA(cudaMalloc(&h->d_outd[rel], kMaxSplitDigits * sizeof(void*)));
        bool any_local = false;
        for (uint32_t d = 0; d < w; ++d) {
            // this chunk's run of digit d starts at cursor == starts[rel][d][chunk]: rebase it to off[d];
            // a null destination keeps that digit in the handle's own split buffer
            const uint64_t start = h->h_shard_starts[((size_t)rel * w + d) * (K + 1) + chunk];
            if (dst[rel][d]) {
                host_ptrs[rel][d] = reinterpret_cast<ulonglong2*>(dst[rel][d]) + (off[rel] ? off[rel][d] : 0) - start;
            } else {
                any_local = true;
                host_ptrs[rel][d] = nullptr;
            }
        }
        if (any_local) {
            int rc = ensure_buffers(h, true, false);
            if (rc != PHJ_OK) return rc;
            for (uint32_t d = 0; d < w; ++d)
                if (!host_ptrs[rel][d]) host_ptrs[rel][d] = h->d_buf_a[rel];
        }
        PHJ_CUDA(cudaMemcpyAsync(h->d_outd[rel], host_ptrs[rel], w * sizeof(void*), cudaMemcpyHostToDevice, h->stream));
        p1.outd[rel] = h->d_outd[rel];
    }
    // segments of this chunk: the build relation's (chunk 0 only) are followed by the probe relation's
    const uint32_t nseg0 = h->nseg1_rel[0];
    const uint32_t first = chunk == 0 ? 0 : nseg0 + h->chunk_first_seg[1][chunk];
    const uint32_t last = nseg0 + h->chunk_first_seg[1][chunk + 1];
    const uint32_t count = last - first;
    if (K > 1) {
        p1.seg_first = first;
        p1.seg_count = count;
    }
    PHJ_CUDA(cudaEventRecord(h->ev[1], h->stream));
    if (count > 0) {
        // split_ctas caps the grid: an NVLink-bound split needs a fraction of the SMs, the rest stay
        // free for the local join of the previous chunk running on another stream
        const uint32_t grid = h->cfg.split_ctas ? std::min<uint32_t>(count, h->cfg.split_ctas) : count;
        KernelScope ks(h, "radix_scatter[split]");
        // remote destinations: the wide-tile kernel; a purely local split keeps the HBM-tuned shape
        if (dst_build && dst_probe && !(h->cfg.flags & (PHJ_FLAG_NO_TMA_STORE | PHJ_FLAG_SPLIT_LOCAL_TILES)))
            PHJ_CUDA(launch_split_scatter(h, h->b1, p1, grid));
        else
            PHJ_CUDA(launch_pass(h, true, h->b1, p1, grid));
    }
    PHJ_CUDA(cudaEventRecord(h->ev[3], h->stream));
    PHJ_CUDA(cudaStreamSynchronize(h->stream));
    PHJ_CUDA(cudaGetLastError());
    out->partition_ns = (uint64_t)((ev_ms(h->ev[1], h->ev[3])) * 1e6);
    out->total_ns = out->partition_ns;
    out->passes = 1;
    out->partitions = h->P;
    out->kernel_launches = h->launches;
    out->hbm_bytes_alg = 16ull * 2 * (h->n[0] + h->n[1]) / K;
    h->launches = 0;
    if (chunk + 1 == K) h->shard_counted = false;
    h->joined_radix = K == 1 && (!dst_build || !dst_probe);
    return PHJ_OK;
}

int phj_shared_alloc(int32_t device, size_t bytes, void** d_ptr, unsigned char* ipc_handle) {
    if (!d_ptr || !ipc_handle) return fail(PHJ_ERR_INVALID, "null argument");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "ipc handle size");
    PHJ_CUDA(cudaSetDevice(device));
    PHJ_CUDA(cudaMalloc(d_ptr, bytes ? bytes : 16));
    cudaIpcMemHandle_t hd;
    cudaError_t e = cudaIpcGetMemHandle(&hd, *d_ptr);
    if (e != cudaSuccess) {
        cudaFree(*d_ptr);
        *d_ptr = nullptr;
        return fail(PHJ_ERR_CUDA, "cudaIpcGetMemHandle failed: %s", cudaGetErrorString(e));
    }
    memcpy(ipc_handle, &hd, 64);
    return PHJ_OK;
}

int phj_enable_peer_access(int32_t device, int32_t peer) {
    PHJ_CUDA(cudaSetDevice(device));
    int can = 0;
    PHJ_CUDA(cudaDeviceCanAccessPeer(&can, device, peer));
    if (!can) return fail(PHJ_ERR_CUDA, "GPU %d cannot access GPU %d's memory", device, peer);
    const cudaError_t e = cudaDeviceEnablePeerAccess(peer, 0);
    if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled)
        return fail(PHJ_ERR_CUDA, "cudaDeviceEnablePeerAccess(%d -> %d) failed: %s", device, peer, cudaGetErrorString(e));
    cudaGetLastError();
    return PHJ_OK;
}

int phj_shared_open(int32_t device, const unsigned char* ipc_handle, void** d_ptr) {
    if (!d_ptr || !ipc_handle) return fail(PHJ_ERR_INVALID, "null argument");
    PHJ_CUDA(cudaSetDevice(device));
    cudaIpcMemHandle_t hd;
    memcpy(&hd, ipc_handle, 64);
    PHJ_CUDA(cudaIpcOpenMemHandle(d_ptr, hd, cudaIpcMemLazyEnablePeerAccess));
    return PHJ_OK;
}

int phj_shared_close(int32_t device, void* d_ptr) {
    PHJ_CUDA(cudaSetDevice(device));
    if (d_ptr) PHJ_CUDA(cudaIpcCloseMemHandle(d_ptr));
    return PHJ_OK;
}

int phj_shared_free(int32_t device, void* d_ptr) {
    PHJ_CUDA(cudaSetDevice(device));
    if (d_ptr) PHJ_CUDA(cudaFree(d_ptr));
    return PHJ_OK;
}

int phj_device_fill_sequential(int32_t device, void* d_out, size_t n, int64_t start) {
    if (n && !d_out) return fail(PHJ_ERR_INVALID, "null argument");
    PHJ_CUDA(cudaSetDevice(device));
    if (n) gen_sequential<<<1184, 256>>>(reinterpret_cast<ulonglong2*>(d_out), n, start);
    PHJ_CUDA(cudaDeviceSynchronize());
    PHJ_CUDA(cudaGetLastError());
    return PHJ_OK;
}

int phj_device_fill_zipf(int32_t device, void* d_out, size_t n, double alpha, int64_t range_first,
                         int64_t range_second, int64_t base_seed, size_t batches) {
    if (n && !d_out) return fail(PHJ_ERR_INVALID, "null argument");
    if (alpha < 0.01) return fail(PHJ_ERR_INVALID, "Zipf: alpha must be >= 0.01");  // Zipf.cpp:18-20
    if (range_first >= range_second) return fail(PHJ_ERR_INVALID, "Zipf: empty range");  // Zipf.cpp:61-67
    if (batches == 0) return fail(PHJ_ERR_INVALID, "Zipf: batches must be > 0");
    if (base_seed <= 0 || (uint64_t)base_seed + batches >= 2147483647ull)
        return fail(PHJ_ERR_INVALID, "Zipf: seeds base_seed .. base_seed + batches must lie in (0, 2^31 - 1)");
    PHJ_CUDA(cudaSetDevice(device));
    ZipfGenParams zp{};
    zp.out = reinterpret_cast<ulonglong2*>(d_out);
    zp.n = n;
    zp.batches = batches;
    zp.batch = n / batches;
    zp.alpha = alpha;
    zp.sd = 1.001 - alpha;
    const double diff = 1.0 - alpha;
    if (std::abs(diff) < 0.01) {
        zp.sd = 0.01 * (diff < 0 ? 1 : -1);
        zp.alpha = 1.0 - zp.sd;
    }
    const uint64_t cardinality = (uint64_t)(range_second - range_first + 1);
    zp.norm = (std::pow((double)cardinality, zp.sd) - zp.alpha) / zp.sd;
    zp.base_seed = base_seed;
    zp.correction = range_first - 1;
    if (n) gen_zipf<<<(uint32_t)((batches + 127) / 128), 128>>>(zp);
    PHJ_CUDA(cudaDeviceSynchronize());
    PHJ_CUDA(cudaGetLastError());
    return PHJ_OK;
}

int phj_memcpy_h2d(int32_t device, void* d_dst, const void* h_src, size_t bytes) {
    if (bytes && (!d_dst || !h_src)) return fail(PHJ_ERR_INVALID, "null argument");
    PHJ_CUDA(cudaSetDevice(device));
    if (bytes) PHJ_CUDA(cudaMemcpy(d_dst, h_src, bytes, cudaMemcpyHostToDevice));
    return PHJ_OK;
}

int phj_memcpy_d2h(int32_t device, void* h_dst, const void* d_src, size_t bytes) {
    if (bytes && (!h_dst || !d_src)) return fail(PHJ_ERR_INVALID, "null argument");
    PHJ_CUDA(cudaSetDevice(device));
    if (bytes) PHJ_CUDA(cudaMemcpy(h_dst, d_src, bytes, cudaMemcpyDeviceToHost));
    return PHJ_OK;
}

int phj_kernel_timing(phj_handle* h, const char* filter) {
    if (!h) return fail(PHJ_ERR_INVALID, "handle is null");
    h->time_kernels = filter != nullptr;
    h->ktime_filter = filter ? filter : "";
    return PHJ_OK;
}

int phj_kernel_times(phj_handle* h, const char** names, uint64_t* ns, uint32_t cap) {
    if (!h) return 0;
    uint32_t n = 0;
    for (int i = 0; i < h->n_ktimes && n < cap; ++i) {
        if (!h->ktimes[i].used) continue;
        names[n] = h->ktimes[i].name;
        ns[n] = (uint64_t)(ev_ms(h->ktimes[i].begin, h->ktimes[i].end) * 1e6);
        ++n;
    }
    return (int)n;
}

uint64_t phj_hash_host(int32_t hash, uint64_t seed, int64_t key) {
    return phj::hash_key_dyn(hash, (uint64_t)key, phj::make_hash_params(hash, seed));
}

int phj_hash_batch(int32_t hash, uint64_t seed, const int64_t* keys, size_t n, uint64_t* out,
                   int32_t device) {
    if (hash < PHJ_HASH_XXH3 || hash > PHJ_HASH_CITY)
        return fail(PHJ_ERR_INVALID, "Unrecognized hash function: %d.", hash);
    if (n == 0) return PHJ_OK;
    if (!keys || !out) return fail(PHJ_ERR_INVALID, "keys/out is null");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        return fail(PHJ_ERR_CUDA, "no CUDA device available; this engine has no CPU fallback");
    }
    PHJ_CUDA(cudaSetDevice(device));
    int64_t* d_keys = nullptr;
    uint64_t* d_out = nullptr;
    PHJ_CUDA(cudaMalloc(&d_keys, n * 8));
    if (cudaMalloc(&d_out, n * 8) != cudaSuccess) {
        cudaFree(d_keys);
        return fail(PHJ_ERR_NOMEM, "cudaMalloc failed");
    }
    cudaMemcpy(d_keys, keys, n * 8, cudaMemcpyHostToDevice);
    const phj::HashParams hp = phj::make_hash_params(hash, seed);
    const uint32_t grid = (uint32_t)((n + 255) / 256);
    switch (hash) {
        case PHJ_HASH_MURMUR3: phj::hash_batch_kernel<phj::kMurmur3><<<grid, 256>>>(d_keys, n, hp, d_out); break;
        case PHJ_HASH_CITY: phj::hash_batch_kernel<phj::kCity><<<grid, 256>>>(d_keys, n, hp, d_out); break;
        default: phj::hash_batch_kernel<phj::kXXH3><<<grid, 256>>>(d_keys, n, hp, d_out); break;
    }
    cudaError_t e = cudaMemcpy(out, d_out, n * 8, cudaMemcpyDeviceToHost);
    cudaFree(d_keys);
    cudaFree(d_out);
    if (e != cudaSuccess) return fail(PHJ_ERR_CUDA, "hash_batch failed: %s", cudaGetErrorString(e));
    return PHJ_OK;
}

int phj_host_alloc(void** out, size_t bytes) {
    if (!out) return fail(PHJ_ERR_INVALID, "out is null");
    PHJ_CUDA(cudaMallocHost(out, bytes ? bytes : 1));
    return PHJ_OK;
}

int phj_host_free(void* p) {
    if (p) PHJ_CUDA(cudaFreeHost(p));
    return PHJ_OK;
}

}  // extern "C"

#include "phj_dist.inl"
