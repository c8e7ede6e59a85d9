// phj_dist.inl -- the radix join sharded over several GPUs; included by phj_engine.cu.
//
// The reference is one process on one machine (no communication backend at all). Radix
// partitions are independent join units, so the join shards by partition with ONE exchange step, and
// that exchange doubles as the (only) partitioning pass:
//
//   count     every rank histograms its row shard by the split digit (owner rank : local partition) = the
//             low log2(world) + b bits of the partitioning hash, PIECE BY PIECE: R and the first probe
//             chunk on stream A, every later chunk on stream C while the pieces before it travel
//             (chunk_scan: 0-based scatter cursors per (digit, piece) + the running chunk starts)
//   sizes     NCCL all-gather of the chunk starts, device to device, once per piece -- the gather that
//             brings the counts of chunk c + 1 is at the same time the barrier behind chunk c; a one-CTA
//             kernel turns them into where each of this rank's runs lands inside its owner's window
//             (digit-major, then source rank: the reference's stable partition order,
//             src/RadixCluster/HashJoin.hpp:394-412), the boundaries of this rank's own partitions, and
//             an overflow flag                                                                [A]
//   scatter   the radix scatter kernel writes every digit run STRAIGHT INTO THE OWNER'S WINDOW over
//             NVLink with its TMA bulk stores -- first R, then S chunk by chunk, each followed by a
//             stream-ordered NCCL collective ("everybody's chunk c has landed")              [A]
//   local     per-partition tables in global memory, hot in L2 (pt_build after R's barrier, pt_probe
//             of chunk c after chunk c's barrier): it runs on stream B WHILE stream A scatters the
//             next chunk -- NVLink-bound stores on one side, L2 / HBM-bound probes on the other
//   reduce    NCCL all-reduce of {matches, overflow flags}; one 32-byte copy to the host
//
// The reference overlaps its two partition pipelines the same way before it waits on either
// (src/RadixCluster/HashJoin.hpp:210-216). In steady state the host enqueues the whole join and waits
// once: sizes never visit the host. Windows and tables are sized by a host-synchronous first pass
// (dist_size) and re-sized only when the device-side overflow flag comes back set.
//
// Two launch modes share this code: one process per GPU (phj_dist_create: NCCL communicator from a
// unique id the caller distributes, windows mapped through CUDA IPC) and one process driving all
// GPUs (phj_config.num_gpus > 1: ncclCommInitAll, one host thread per GPU, peer access).
// NCCL is bound at run time (dlopen), so the single-GPU library has no NCCL dependency.
#include <dlfcn.h>
#include <nccl.h>
#include <pthread.h>

#include "phj_dist_kernels.cuh"

using phj::kMaxHot;
using phj::LayoutParams;
using phj::PullDesc;

#include <condition_variable>
#include <functional>
#include <mutex>

namespace {

constexpr int kMaxRanks = 16;

struct NcclApi {
    void* lib = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommInitAll)(ncclComm_t*, int, const int*) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    ncclResult_t (*GetVersion)(int*) = nullptr;
    std::string error;
};

NcclApi* nccl_api() {
    static NcclApi api;
    static std::once_flag once;
    std::call_once(once, [] {
        // inside a torch process this resolves to the NCCL torch has already loaded (same SONAME)
        api.lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_LOCAL);
        if (!api.lib) api.lib = dlopen("libnccl.so", RTLD_NOW | RTLD_LOCAL);
        if (!api.lib) {
            api.error = std::string("libnccl.so.2 not found: ") + dlerror();
            return;
        }
        auto sym = [&](const char* name) {
            void* p = dlsym(api.lib, name);
            if (!p) api.error = std::string("NCCL symbol missing: ") + name;
            return p;
        };
        api.GetUniqueId = reinterpret_cast<decltype(api.GetUniqueId)>(sym("ncclGetUniqueId"));
        api.CommInitRank = reinterpret_cast<decltype(api.CommInitRank)>(sym("ncclCommInitRank"));
        api.CommInitAll = reinterpret_cast<decltype(api.CommInitAll)>(sym("ncclCommInitAll"));
        api.CommDestroy = reinterpret_cast<decltype(api.CommDestroy)>(sym("ncclCommDestroy"));
        api.AllGather = reinterpret_cast<decltype(api.AllGather)>(sym("ncclAllGather"));
        api.AllReduce = reinterpret_cast<decltype(api.AllReduce)>(sym("ncclAllReduce"));
        api.GetErrorString = reinterpret_cast<decltype(api.GetErrorString)>(sym("ncclGetErrorString"));
        api.GetVersion = reinterpret_cast<decltype(api.GetVersion)>(sym("ncclGetVersion"));
    });
    return api.error.empty() ? &api : nullptr;
}

#define PHJ_NCCL(call)                                                                              \
    do {                                                                                            \
        ncclResult_t r_ = (call);                                                                   \
        if (r_ != ncclSuccess)                                                                      \
            return fail(PHJ_ERR_CUDA, "%s failed: %s (%s:%d)", #call, nccl_api()->GetErrorString(r_), \
                        __FILE__, __LINE__);                                                        \
    } while (0)

}  // namespace

// ---- one rank -------------------------------------------------------------------------------------
struct phj_group;

struct phj_dist {
    int rank = 0, world = 1, device = 0;
    phj_config cfg{};
    phj_group* group = nullptr;  // set when all ranks live in this process
    ncclComm_t comm = nullptr;
    bool own_comm = false;
    phj_handle* split = nullptr;  // PHJ_ALGO_SHARD_SPLIT handle: stream A, histogram, scan, scatter
    cudaStream_t sb = nullptr;    // stream B: the local join
    cudaStream_t sc = nullptr;    // stream C: the counts of the probe chunks behind the first
    uint32_t b_local = 0, d_local = 1, ndig = 1, K = 1;
    // windows: 0 = build, 1 = probe
    void* win[2] = {nullptr, nullptr};
    unsigned char win_handle[2][64] = {};
    void* peer[2][kMaxRanks] = {};
    uint64_t caps[2][kMaxRanks] = {};  // tuples; tracked identically by every rank
    // device-side plan
    uint64_t* d_all_starts = nullptr;
    uint64_t* h_all_starts = nullptr;  // pinned
    ulonglong2** d_peer = nullptr;     // [2][kMaxRanks]
    ulonglong2** d_outd = nullptr;     // [(1 + K)][ndig]
    uint64_t* d_lb = nullptr;          // build [d_local + 1], then probe [K][d_local + 1]
    unsigned long long* d_flags = nullptr;  // [4] local, [4..8) all-reduced
    unsigned long long* h_flags = nullptr;  // pinned [8]
    unsigned char* d_xchg = nullptr;        // handle exchange staging: [(1 + world)][128]
    unsigned char* h_xchg = nullptr;        // pinned
    // L2 tables
    uint64_t* d_pt = nullptr;
    size_t cap_pt = 0;
    uint32_t region_buckets = 0, max_keys = 0;
    uint32_t* d_ptflags = nullptr;
    cudaEvent_t ev_r = nullptr, ev_c[kMaxSplitChunks] = {}, ev_cnt[kMaxSplitChunks] = {}, ev_local = nullptr, ev_t[8] = {};
    bool sized = false, have_data = false;
    bool piecewise = false;  // count the probe chunks behind the first while the pieces before them travel
    // heavy hitters: decided by the sizing pass, identically on every rank
    uint32_t n_hot = 0, hot[kMaxHot] = {};
    uint32_t np = 1;               // partitions of this rank: d_local + the hot digits of other owners
    PullDesc* d_pulls = nullptr;
    size_t n[2] = {0, 0};
    uint64_t sent_remote_bytes = 0;  // this rank's tuples that leave the GPU, from the last sizing pass
    uint32_t resizes = 0;
};

struct phj_group {
    int world = 0;
    pthread_barrier_t bar;
    bool bar_ok = false;
    std::vector<phj_dist*> ranks;
    void* win[2][kMaxRanks] = {};
    struct Worker {
        std::thread thread;
        std::mutex m;
        std::condition_variable cv;
        std::function<int()> job;
        bool has_job = false, done = false, stop = false;
        int rc = 0;
        std::string error;
    };
    std::vector<Worker*> workers;
    // sharding of the host relations given to phj_upload
    std::vector<ncclComm_t> comms;
    // NO_PARTITIONING over several GPUs needs no exchange: every GPU gets the whole build relation
    // and a row shard of the probe relation, builds its own table and probes its shard; the counts add up.
    // One ordinary single-GPU handle per device, no NCCL, no peer access.
    std::vector<phj_handle*> replicas;
};

namespace {

int dist_host_barrier(phj_dist* D) {
    // every rank's stream A has drained AND every rank has arrived
    NcclApi* nc = nccl_api();
    PHJ_NCCL(nc->AllReduce(D->d_flags + 8, D->d_flags + 8, 1, ncclUint64, ncclSum, D->comm, D->split->stream));
    PHJ_CUDA(cudaStreamSynchronize(D->split->stream));
    return PHJ_OK;
}

// Publish this rank's window pointers and map everybody else's.
int dist_exchange_windows(phj_dist* D) {
    NcclApi* nc = nccl_api();
    const int W = D->world;
    if (D->group) {
        for (int w = 0; w < 2; ++w) D->group->win[w][D->rank] = D->win[w];
        pthread_barrier_wait(&D->group->bar);
        for (int w = 0; w < 2; ++w)
            for (int r = 0; r < W; ++r) D->peer[w][r] = D->group->win[w][r];
        pthread_barrier_wait(&D->group->bar);
    } else {
        memcpy(D->h_xchg, D->win_handle[0], 64);
        memcpy(D->h_xchg + 64, D->win_handle[1], 64);
        PHJ_CUDA(cudaMemcpyAsync(D->d_xchg, D->h_xchg, 128, cudaMemcpyHostToDevice, D->split->stream));
        PHJ_NCCL(nc->AllGather(D->d_xchg, D->d_xchg + 128, 128, ncclUint8, D->comm, D->split->stream));
        PHJ_CUDA(cudaMemcpyAsync(D->h_xchg + 128, D->d_xchg + 128, (size_t)128 * W, cudaMemcpyDeviceToHost,
                                 D->split->stream));
        PHJ_CUDA(cudaStreamSynchronize(D->split->stream));
        for (int r = 0; r < W; ++r)
            for (int w = 0; w < 2; ++w) {
                if (r == D->rank) {
                    D->peer[w][r] = D->win[w];
                    continue;
                }
                cudaIpcMemHandle_t hd;
                memcpy(&hd, D->h_xchg + 128 + (size_t)128 * r + 64 * w, 64);
                PHJ_CUDA(cudaIpcOpenMemHandle(&D->peer[w][r], hd, cudaIpcMemLazyEnablePeerAccess));
            }
    }
    ulonglong2* table[2][kMaxRanks] = {};
    for (int w = 0; w < 2; ++w)
        for (int r = 0; r < W; ++r) table[w][r] = reinterpret_cast<ulonglong2*>(D->peer[w][r]);
    PHJ_CUDA(cudaMemcpyAsync(D->d_peer, table, sizeof(table), cudaMemcpyHostToDevice, D->split->stream));
    PHJ_CUDA(cudaStreamSynchronize(D->split->stream));
    return PHJ_OK;
}

int dist_close_peers(phj_dist* D) {
    for (int w = 0; w < 2; ++w)
        for (int r = 0; r < D->world; ++r) {
            if (D->peer[w][r] && r != D->rank && !D->group) PHJ_CUDA(cudaIpcCloseMemHandle(D->peer[w][r]));
            D->peer[w][r] = nullptr;
        }
    return PHJ_OK;
}

// Counting pieces [first, last] of the row shard (0: R; 1 + c: chunk c of S) on `stream`: ONE histogram launch
// over their segments by split digit (the pieces of a relation are consecutive segments; R and S are adjacent too),
// then chunk_scan: every piece's 0-based scatter cursors and its per-digit sizes.
int dist_enqueue_count_pieces(phj_dist* D, uint32_t first, uint32_t last, cudaStream_t stream) {
    phj_handle* h = D->split;
    const uint32_t K = D->K;
    if (!h->d_shard_starts) {
        const size_t cap = (size_t)2 * kMaxSplitDigits * (kMaxSplitChunks + 1);
        PHJ_CUDA(cudaMalloc(&h->d_shard_starts, cap * 8));
        PHJ_CUDA(cudaMemset(h->d_shard_starts, 0, cap * 8));
        PHJ_CUDA(cudaMallocHost(&h->h_shard_starts, cap * 8));
    }
    const uint32_t nseg0 = h->nseg1_rel[0];
    ChunkScanParams cs{};
    cs.counts = h->d_counts;
    cs.cursors = h->d_cursors;
    cs.sizes = h->d_shard_starts;
    for (int rel = 0; rel < 2; ++rel) {
        cs.cnt_base[rel] = h->cnt_base1_rel[rel];
        cs.nseg_rel[rel] = h->nseg1_rel[rel];
    }
    cs.ndig = D->ndig;
    cs.K = K;
    cs.piece_first = first;
    cs.seg_first[0] = 0;
    cs.seg_count[0] = nseg0;
    for (uint32_t c = 0; c < K; ++c) {
        cs.seg_first[1 + c] = h->chunk_first_seg[1][c];
        cs.seg_count[1 + c] = h->chunk_first_seg[1][c + 1] - h->chunk_first_seg[1][c];
    }
    // segments of the pieces in the handle's segment list: R's, then S's
    const uint32_t seg_lo = first ? nseg0 + cs.seg_first[first] : 0;
    const uint32_t seg_hi = last ? nseg0 + cs.seg_first[last] + cs.seg_count[last] : nseg0;
    cudaStream_t keep = h->stream;
    h->stream = stream;  // launch_pass and KernelScope follow the handle's stream
    cudaError_t e = cudaSuccess;
    if (seg_hi > seg_lo) {
        PassParams p1{};
        fill_pass1_params(h, p1);
        p1.seg_first = seg_lo;
        p1.seg_count = seg_hi - seg_lo;
        KernelScope ks(h, "radix_histogram[split]");
        e = launch_pass(h, false, h->b1, p1, seg_hi - seg_lo);
    }
    if (e == cudaSuccess) {
        KernelScope ks(h, "chunk_scan");
        chunk_scan<<<dim3(D->ndig, last - first + 1), 256, 0, stream>>>(cs);
        e = cudaGetLastError();
    }
    h->stream = keep;
    PHJ_CUDA(e);
    return PHJ_OK;
}

// All-gather of the per-piece sizes on stream A (5 KB per rank; entries of pieces that are still being counted are
// in flux and not read by anybody yet). Stream-ordered, so it is also a barrier: when it completes HERE, every rank
// has finished what it had enqueued before it.
int dist_enqueue_gather(phj_dist* D) {
    NcclApi* nc = nccl_api();
    phj_handle* h = D->split;
    const size_t per_rank = (size_t)2 * D->ndig * (D->K + 1);
    PHJ_NCCL(nc->AllGather(h->d_shard_starts, D->d_all_starts, per_rank, ncclUint64, D->comm, h->stream));
    h->launches += 1;
    return PHJ_OK;
}

// The whole count at once on stream A.
int dist_enqueue_count(phj_dist* D) {
    int rc = dist_enqueue_count_pieces(D, 0, D->K, D->split->stream);
    if (rc != PHJ_OK) return rc;
    return dist_enqueue_gather(D);
}

// Host-synchronous sizing pass (first join, new relations, or after an overflow): counts as the join
// itself does, then every rank derives -- from the same all-gathered numbers, hence identically --
// how large every rank's windows and tables must be, and the ranks whose windows are too small
// re-allocate them; all windows are then (re)mapped.
int dist_size(phj_dist* D) {
    phj_handle* h = D->split;
    const int W = D->world;
    const uint32_t K = D->K, ndig = D->ndig, dl = D->d_local;
    int rc = dist_enqueue_count(D);
    if (rc != PHJ_OK) return rc;
    const size_t total = (size_t)W * 2 * ndig * (K + 1);
    PHJ_CUDA(cudaMemcpyAsync(D->h_all_starts, D->d_all_starts, total * 8, cudaMemcpyDeviceToHost, h->stream));
    PHJ_CUDA(cudaStreamSynchronize(h->stream));
    PHJ_CUDA(cudaGetLastError());
    // per (source, relation, digit): tuples over all pieces (R has one piece, entry 0; S one per chunk)
    auto size_of = [&](int src, int rel, uint32_t d) {
        const uint64_t* st = D->h_all_starts + (((size_t)src * 2 + rel) * ndig + d) * (K + 1);
        uint64_t n = 0;
        for (uint32_t c = 0; c < (rel ? K : 1); ++c) n += st[c];
        return n;
    };
    std::vector<uint64_t> tot_r(ndig, 0), tot_s(ndig, 0);
    uint64_t total_s = 0;
    for (uint32_t d = 0; d < ndig; ++d)
        for (int src = 0; src < W; ++src) {
            tot_r[d] += size_of(src, 0, d);
            tot_s[d] += size_of(src, 1, d);
        }
    for (uint32_t d = 0; d < ndig; ++d) total_s += tot_s[d];
    // Heavy hitters: split digits whose probe side alone outweighs a quarter of one rank's fair share. Never the case
    // for uniform keys (a digit is world / digits of a share); key 1 at Zipf 1.25 is 22 % of S. Every rank derives the
    // same set from the same all-gathered sizes.
    D->n_hot = 0;
    if (!(D->cfg.flags & PHJ_FLAG_NO_HOT_DIGITS) && W > 1 && ndig > (uint32_t)W)
        for (uint32_t d = 0; d < ndig && D->n_hot < (uint32_t)kMaxHot; ++d)
            if (tot_s[d] * (uint64_t)W * 4 > std::max<uint64_t>(total_s, 1)) D->hot[D->n_hot++] = d;
    std::vector<bool> is_hot(ndig, false);
    for (uint32_t i = 0; i < D->n_hot; ++i) is_hot[D->hot[i]] = true;
    uint64_t need[2][kMaxRanks] = {};
    uint64_t max_part = 0, mine_remote = 0;
    for (uint32_t d = 0; d < ndig; ++d) {
        const int owner = (int)(d / dl);
        max_part = std::max(max_part, tot_r[d]);
        if (owner != D->rank) mine_remote += size_of(D->rank, 0, d) + (is_hot[d] ? 0 : size_of(D->rank, 1, d));
        for (int r = 0; r < W; ++r) {
            if (r == owner) {
                need[0][r] += tot_r[d];
                need[1][r] += is_hot[d] ? size_of(r, 1, d) : tot_s[d];
            } else if (is_hot[d]) {  // an extra partition: the build side pulled from the owner, the own probe tuples
                need[0][r] += tot_r[d];
                need[1][r] += size_of(r, 1, d);
            }
        }
    }
    D->np = dl;
    for (uint32_t i = 0; i < D->n_hot; ++i)
        if ((int)(D->hot[i] / dl) != D->rank) ++D->np;
    D->sent_remote_bytes = 16 * mine_remote;
    bool grow_any = false, grow_mine[2] = {false, false};
    for (int w = 0; w < 2; ++w)
        for (int r = 0; r < W; ++r)
            if (need[w][r] > D->caps[w][r] || (D->caps[w][r] == 0)) {
                D->caps[w][r] = need[w][r] + need[w][r] / 8 + 4096;
                grow_any = true;
                if (r == D->rank) grow_mine[w] = true;
            }
    // An allocation that fails on ONE GPU must not leave the others waiting in the next collective: failures are
    // recorded, all ranks agree on them (one word, all-reduced), and either everybody goes on or everybody returns.
    int alloc_rc = PHJ_OK;
    auto device_alloc = [&](void** p, size_t bytes, const char* what) {
        *p = nullptr;
        const cudaError_t e = cudaMalloc(p, bytes);
        if (e == cudaSuccess) return true;
        cudaGetLastError();
        *p = nullptr;
        alloc_rc = fail(PHJ_ERR_NOMEM, "GPU %d: no memory for %s of the sharded join (%zu bytes): %s", D->device, what,
                        bytes, cudaGetErrorString(e));
        return false;
    };
    const bool remap = grow_any || !D->peer[0][D->rank];
    if (remap) {
        if ((rc = dist_close_peers(D)) != PHJ_OK) return rc;
        if ((rc = dist_host_barrier(D)) != PHJ_OK) return rc;  // nobody maps a window that is about to go
        for (int w = 0; w < 2; ++w)
            if (grow_mine[w] || !D->win[w]) {
                if (D->win[w]) PHJ_CUDA(cudaFree(D->win[w]));
                D->win[w] = nullptr;
                if (device_alloc(&D->win[w], D->caps[w][D->rank] * 16, w ? "the probe window" : "the build window") &&
                    !D->group) {
                    cudaIpcMemHandle_t hd;
                    PHJ_CUDA(cudaIpcGetMemHandle(&hd, D->win[w]));
                    memcpy(D->win_handle[w], &hd, 64);
                }
            }
    }
    // tables: load <= 0.5 at the largest build partition seen (duplicates only lower it), 10 % to spare
    uint32_t rb = 64;
    while ((uint64_t)rb * 2 < max_part + max_part / 10) rb <<= 1;
    {
        const size_t want = (size_t)D->np * rb * 4;  // one region per partition of this rank
        if (want > D->cap_pt || !D->d_pt) {
            if (D->d_pt) PHJ_CUDA(cudaFree(D->d_pt));
            D->d_pt = nullptr;
            D->cap_pt = 0;
            if (device_alloc(reinterpret_cast<void**>(&D->d_pt), want * 8, "the partition tables")) D->cap_pt = want;
        }
        D->region_buckets = rb;
    }
    {
        NcclApi* nc = nccl_api();
        D->h_flags[0] = alloc_rc != PHJ_OK;
        PHJ_CUDA(cudaMemcpyAsync(D->d_flags + 9, D->h_flags, 8, cudaMemcpyHostToDevice, h->stream));
        PHJ_NCCL(nc->AllReduce(D->d_flags + 9, D->d_flags + 9, 1, ncclUint64, ncclSum, D->comm, h->stream));
        PHJ_CUDA(cudaMemcpyAsync(D->h_flags, D->d_flags + 9, 8, cudaMemcpyDeviceToHost, h->stream));
        PHJ_CUDA(cudaStreamSynchronize(h->stream));
        if (D->h_flags[0]) {
            if (alloc_rc != PHJ_OK) return alloc_rc;
            return fail(PHJ_ERR_NOMEM, "sharded join: %llu of the %d GPUs could not allocate their windows or tables",
                        (unsigned long long)D->h_flags[0], W);
        }
    }
    if (remap) {
        if ((rc = dist_exchange_windows(D)) != PHJ_OK) return rc;
        ++D->resizes;
    }
    D->max_keys = D->region_buckets * 3;
    D->sized = true;
    return PHJ_OK;
}

// The whole join, enqueued.
int dist_enqueue_join(phj_dist* D) {
    NcclApi* nc = nccl_api();
    phj_handle* h = D->split;
    cudaStream_t sa = h->stream, sb = D->sb, sc = D->sc;
    const uint32_t K = D->K, ndig = D->ndig, dl = D->d_local;
    int rc;
    PHJ_CUDA(cudaEventRecord(D->ev_t[0], sa));
    PHJ_CUDA(cudaMemsetAsync(D->d_flags, 0, 4 * sizeof(unsigned long long), sa));
    PHJ_CUDA(cudaMemsetAsync(D->d_ptflags, 0, 4, sa));
    // ---- count. Up front (the default): the whole row shard on stream A, one gather, one layout launch.
    //      Piece-wise (PHJ_FLAG_COUNT_PIECEWISE): R and the first probe chunk on stream A, the other chunks on
    //      stream C while the pieces before them travel, the gather of chunk c + 1's sizes doubling as the barrier
    //      behind chunk c: 0.3 ms less exposed counting, paid back by a slower shuffle (the histograms take SMs
    //      and HBM from the scatter: 569 instead of 590 GB/s) and 70 us more per chunk boundary -- 5.81 against
    //      5.87 ms at 8 GPUs; at 2 GPUs the GPU is busy throughout and it loses (profiles/r02_multigpu.md). ----
    const bool piecewise = D->piecewise && K > 1;
    if ((rc = dist_enqueue_count_pieces(D, 0, piecewise ? 1 : K, sa)) != PHJ_OK) return rc;
    if (piecewise) {
        PHJ_CUDA(cudaEventRecord(D->ev_cnt[0], sa));
        PHJ_CUDA(cudaStreamWaitEvent(sc, D->ev_cnt[0], 0));
        for (uint32_t c = 1; c < K; ++c) {
            if ((rc = dist_enqueue_count_pieces(D, 1 + c, 1 + c, sc)) != PHJ_OK) return rc;
            PHJ_CUDA(cudaEventRecord(D->ev_cnt[c], sc));
        }
    }
    if ((rc = dist_enqueue_gather(D)) != PHJ_OK) return rc;
    PHJ_CUDA(cudaEventRecord(D->ev_t[1], sa));
    LayoutParams lp{};
    lp.all_sizes = D->d_all_starts;
    lp.world = (uint32_t)D->world;
    lp.rank = (uint32_t)D->rank;
    lp.ndig = ndig;
    lp.d_local = dl;
    lp.K = K;
    lp.peer_build = D->d_peer;
    lp.peer_probe = D->d_peer + kMaxRanks;
    lp.outd = D->d_outd;
    const uint32_t np = D->np;  // this rank's partitions: the owned ones + the hot digits of other owners
    lp.lb_build = D->d_lb;
    lp.lb_probe = D->d_lb + (np + 1);
    lp.np = np;
    lp.n_hot = D->n_hot;
    for (uint32_t i = 0; i < D->n_hot; ++i) lp.hot[i] = D->hot[i];
    lp.pulls = D->d_pulls;
    lp.cap_build = D->caps[0][D->rank];
    lp.cap_probe = D->caps[1][D->rank];
    lp.max_keys = D->max_keys;
    lp.flags = D->d_flags;
    auto layout = [&](uint32_t first, uint32_t last) {
        lp.which_first = first;
        lp.which_last = last;
        KernelScope ks(h, "dist_layout");
        dist_layout<<<1, 256, 0, sa>>>(lp);
    };
    layout(0, piecewise ? 1 : K);
    // ---- stream B: clear the tables while stream A still scatters R ----
    PHJ_CUDA(cudaEventRecord(D->ev_t[2], sa));
    PHJ_CUDA(cudaStreamWaitEvent(sb, D->ev_t[2], 0));  // the build boundaries are ready, the flags are zero
    // The two legs must CO-RESIDE to overlap: a scatter CTA (1024 threads x 64 registers) owns a whole SM's
    // register file, as do six probe CTAs, so left to themselves the two kernels only take turns on the SMs
    // (measured at 2 GPUs: 5.4 ms with 4 chunks against 4.5 ms without any overlap). So the SMs are split. The
    // probe -- whose tiles are dealt round-robin to however many CTAs there are -- gets as many SMs as it needs
    // to finish a chunk while the next chunk's scatter is bound by NVLink ((W - 1) / W of the chunk at ~0.6 TB/s
    // against ~0.3 ms of a whole GPU for the probe), the scatter the rest: more scatter CTAs mean more stores in
    // flight (2 GPUs, whole GPU: 660 GB/s; 80 CTAs: 560 GB/s). Measured at 2 GPUs, scatter / probe SMs:
    // 80 / 68: 4.07 ms; 100 / 48: 4.28; 110 / 38: 4.75; 120 / 28: 5.42. With one chunk nothing overlaps and
    // both kernels get the whole GPU.
    uint32_t grid_l = (uint32_t)h->sm_count * 8, scatter_cap = 0;
    if (K > 1) {
        const uint32_t sms = (uint32_t)h->sm_count;
        const double remote = (double)(D->world - 1) / D->world;
        uint32_t probe_sms = remote > 0 ? (uint32_t)(sms * 0.30 / (1.29 * remote) + 0.5) : sms / 2;
        probe_sms = std::min(std::max(probe_sms, sms / 5), sms / 2);
        scatter_cap = D->cfg.split_ctas ? std::min<uint32_t>(D->cfg.split_ctas, sms - 1) : sms - probe_sms;
        scatter_cap = std::max<uint32_t>(scatter_cap, 1);
        grid_l = std::max<uint32_t>(sms - scatter_cap, 1) * 6;  // pt_probe: 256 threads x <= 40 registers = 6 CTAs per SM
    }
    {
        KernelScope ks(h, "pt_clear", 1, sb);
        gt_clear<<<grid_l, 256, 0, sb>>>(D->d_pt, (uint64_t)np * D->region_buckets * 4);
    }

    // ---- stream A: R, then S chunk by chunk, into the owners' windows ----
    PassParams p1{};
    fill_pass1_params(h, p1);
    p1.cursor_bias[0] = p1.cursor_bias[1] = 0;  // chunk_scan's cursors are 0-based per (digit, piece)
    const uint32_t nseg0 = h->nseg1_rel[0];
    auto scatter = [&](uint32_t first, uint32_t count, uint32_t table) -> int {
        PassParams pp = p1;
        pp.seg_first = first;
        pp.seg_count = count;
        pp.outd[0] = pp.outd[1] = D->d_outd + (size_t)table * ndig;
        if (count) {
            KernelScope ks(h, table == 0 ? "radix_scatter[shuffle R]" : "radix_scatter[shuffle S]");
            PHJ_CUDA(launch_split_scatter(h, h->b1, pp, scatter_cap ? std::min(count, scatter_cap) : count));
        }
        return PHJ_OK;
    };
    // stream-ordered barrier: when it completes HERE, every rank's scatter of this piece has completed
    auto barrier = [&]() -> int {
        PHJ_NCCL(nc->AllReduce(D->d_flags + 8, D->d_flags + 8, 1, ncclUint64, ncclSum, D->comm, sa));
        h->launches += 1;
        return PHJ_OK;
    };
    if ((rc = scatter(0, nseg0, 0)) != PHJ_OK) return rc;
    if ((rc = barrier()) != PHJ_OK) return rc;
    if (np > dl) {
        // heavy hitters: their build partitions, complete in the owners' windows now, are copied here
        KernelScope ks(h, "dist_pull");
        dist_pull<<<dim3(64, np - dl), 256, 0, sa>>>(D->d_pulls);
    }
    PHJ_CUDA(cudaEventRecord(D->ev_r, sa));

    PtParams q{};
    q.build = reinterpret_cast<const ulonglong2*>(D->win[0]);
    q.probe = reinterpret_cast<const ulonglong2*>(D->win[1]);
    q.bounds_build = D->d_lb;
    q.npart = np;
    q.max_keys = D->max_keys;
    q.table_mul = (h->cfg.table_seed * 0x9E3779B97F4A7C15ULL) | 1ULL;
    if (h->cfg.table_seed == 0) q.table_mul = 0xBF58476D1CE4E5B9ULL;
    q.table_mul |= 1ULL << 32;
    {
        const HashParams hp = make_hash_params(h->cfg.hash, h->cfg.hash_seed);
        const uint64_t hs = hash_key_dyn(h->cfg.hash, kEmptyKey, hp);
        const uint64_t digit = (ndig & (ndig - 1)) ? hs % ndig : hs & (ndig - 1);
        q.sentinel_part = digit / dl == (uint64_t)D->rank ? (uint32_t)(digit % dl) : 0xffffffffu;
        uint32_t extra = dl;  // ... or one of the hot digits of other owners kept here
        for (uint32_t i = 0; i < D->n_hot; ++i) {
            if (D->hot[i] / dl == (uint32_t)D->rank) continue;
            if (D->hot[i] == digit) q.sentinel_part = extra;
            ++extra;
        }
    }
    q.table = D->d_pt;
    q.region_buckets = D->region_buckets;
    q.region_shift32 = 32 - (uint32_t)ilog2_ceil(D->region_buckets);
    q.flags = D->d_ptflags;
    q.matches = D->d_flags;
    PHJ_CUDA(cudaStreamWaitEvent(sb, D->ev_r, 0));
    q.bounds_probe = D->d_lb + (np + 1);
    {
        KernelScope ks(h, "pt_build", 1, sb);
        pt_build<256><<<grid_l, 256, 0, sb>>>(q);
    }
    PHJ_CUDA(cudaEventRecord(D->ev_t[5], sb));

    for (uint32_t c = 0; c < K; ++c) {
        const uint32_t first = nseg0 + h->chunk_first_seg[1][c], last = nseg0 + h->chunk_first_seg[1][c + 1];
        if ((rc = scatter(first, last - first, 1 + c)) != PHJ_OK) return rc;
        if (piecewise && c + 1 < K) {
            // the all-gather that brings everybody's counts of chunk c + 1 is also the barrier behind chunk c
            PHJ_CUDA(cudaStreamWaitEvent(sa, D->ev_cnt[c + 1], 0));
            if ((rc = dist_enqueue_gather(D)) != PHJ_OK) return rc;
        } else if ((rc = barrier()) != PHJ_OK) {
            return rc;
        }
        // The probe of chunk c is released right in front of the launch of the next scatter, never earlier: its
        // CTAs spread over every free SM, and a scatter CTA needs a whole one (measured at 8 GPUs with the layout
        // kernel in between: 1.76 instead of 1.16 ms per chunk).
        if (piecewise && c + 1 < K) layout(2 + c, 2 + c);
        PHJ_CUDA(cudaEventRecord(D->ev_c[c], sa));
        PHJ_CUDA(cudaStreamWaitEvent(sb, D->ev_c[c], 0));
        q.bounds_probe = D->d_lb + (np + 1) + (size_t)c * (np + 1);
        {
            // the last chunk's probe runs after the last scatter: it gets the whole GPU
            const uint32_t grid_p = c + 1 == K ? (uint32_t)h->sm_count * 6 : grid_l;  // one wave: six CTAs per SM
            KernelScope ks(h, "pt_probe", 1, sb);
            pt_probe<256><<<grid_p, 256, 0, sb>>>(q);
        }
    }
    PHJ_CUDA(cudaEventRecord(D->ev_t[3], sa));  // the shuffle is complete
    PHJ_CUDA(cudaEventRecord(D->ev_local, sb));
    PHJ_CUDA(cudaStreamWaitEvent(sa, D->ev_local, 0));
    // {matches, overflow, oversize}: summed over the ranks; also the barrier that keeps the next join's
    // stores out of windows that are still being probed
    PHJ_NCCL(nc->AllReduce(D->d_flags, D->d_flags + 4, 4, ncclUint64, ncclSum, D->comm, sa));
    h->launches += 1;
    PHJ_CUDA(cudaMemcpyAsync(D->h_flags, D->d_flags, 8 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, sa));
    PHJ_CUDA(cudaEventRecord(D->ev_t[4], sa));
    return PHJ_OK;
}

int dist_join_rank(phj_dist* D, phj_result* out) {
    if (!D->have_data) return fail(PHJ_ERR_STATE, "join called before upload / bind_device");
    PHJ_CUDA(cudaSetDevice(D->device));
    phj_handle* h = D->split;
    memset(out, 0, sizeof(*out));
    h->launches = 0;
    h->n_ktimes = 0;
    int rc;
    for (int attempt = 0;; ++attempt) {
        // the sizing pass counts on its own; the join below counts again, so that every join runs -- and
        // times -- the same work
        if (!D->sized && (rc = dist_size(D)) != PHJ_OK) return rc;
        h->n_ktimes = 0;
        if ((rc = dist_enqueue_join(D)) != PHJ_OK) return rc;
        PHJ_CUDA(cudaStreamSynchronize(h->stream));
        PHJ_CUDA(cudaStreamSynchronize(D->sb));
        PHJ_CUDA(cudaStreamSynchronize(D->sc));
        PHJ_CUDA(cudaGetLastError());
        if (D->h_flags[5] == 0 && D->h_flags[6] == 0) break;  // no window / table anywhere was too small
        if (attempt == 1)
            return fail(PHJ_ERR_INVALID, "sharded join: windows or tables still too small after re-sizing "
                                         "(%llu window overflows, %llu oversize partitions)",
                        (unsigned long long)D->h_flags[5], (unsigned long long)D->h_flags[6]);
        D->sized = false;  // every rank saw the same all-reduced flags: all of them re-size
    }
    out->matches = D->h_flags[4];
    out->partition_ns = (uint64_t)(ev_ms(D->ev_t[0], D->ev_t[3]) * 1e6);
    out->count_ns = (uint64_t)(ev_ms(D->ev_t[0], D->ev_t[1]) * 1e6);
    out->shuffle_ns = (uint64_t)(ev_ms(D->ev_t[2], D->ev_t[3]) * 1e6);
    out->join_ns = (uint64_t)(ev_ms(D->ev_r, D->ev_local) * 1e6);  // tables + probes (overlapping the shuffle)
    out->build_ns = (uint64_t)(ev_ms(D->ev_r, D->ev_t[5]) * 1e6);  // table build, once the build side has landed
    out->probe_ns = (uint64_t)(ev_ms(D->ev_t[5], D->ev_local) * 1e6);  // the probes of all chunks (incl. waiting for them)
    out->total_ns = (uint64_t)(ev_ms(D->ev_t[0], D->ev_t[4]) * 1e6);
    out->passes = 1;
    out->partitions = D->ndig;
    out->gpus = (uint32_t)D->world;
    out->shuffle_bytes = D->sent_remote_bytes;
    out->kernel_launches = h->launches;
    out->d2h_bytes = 8 * sizeof(unsigned long long);
    // per rank: histogram read + scatter read + remote / local write + probe read
    out->hbm_bytes_alg = 16ull * 4 * (D->n[0] + D->n[1]);
    return PHJ_OK;
}

int dist_alloc(phj_dist* D) {
    PHJ_CUDA(cudaSetDevice(D->device));
    const size_t total = (size_t)D->world * 2 * D->ndig * (D->K + 1);
    PHJ_CUDA(cudaMalloc(&D->d_all_starts, total * 8));
    PHJ_CUDA(cudaMallocHost(&D->h_all_starts, total * 8));
    PHJ_CUDA(cudaMalloc(&D->d_peer, sizeof(ulonglong2*) * 2 * kMaxRanks));
    PHJ_CUDA(cudaMalloc(&D->d_outd, sizeof(ulonglong2*) * (size_t)(1 + D->K) * D->ndig));
    PHJ_CUDA(cudaMalloc(&D->d_lb, 8 * (size_t)(1 + D->K) * (D->d_local + kMaxHot + 1)));
    PHJ_CUDA(cudaMalloc(&D->d_pulls, sizeof(PullDesc) * kMaxHot));
    PHJ_CUDA(cudaMalloc(&D->d_flags, 16 * sizeof(unsigned long long)));
    PHJ_CUDA(cudaMemset(D->d_flags, 0, 16 * sizeof(unsigned long long)));
    PHJ_CUDA(cudaMallocHost(&D->h_flags, 8 * sizeof(unsigned long long)));
    PHJ_CUDA(cudaMalloc(&D->d_xchg, (size_t)128 * (1 + D->world)));
    PHJ_CUDA(cudaMallocHost(&D->h_xchg, (size_t)128 * (1 + D->world)));
    PHJ_CUDA(cudaMalloc(&D->d_ptflags, 16));
    // CTAs of stream A (the scatter, highest priority) go first whenever an SM slot frees up, then the histograms
    // of the later chunks (C: a late count would stall the NVLink scatter behind it), then the local join (B)
    int prio_lo = 0, prio_hi = 0;
    PHJ_CUDA(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
    PHJ_CUDA(cudaStreamCreateWithPriority(&D->sb, cudaStreamNonBlocking, prio_lo));
    PHJ_CUDA(cudaStreamCreateWithPriority(&D->sc, cudaStreamNonBlocking, (prio_lo + prio_hi) / 2));
    for (auto& e : D->ev_cnt) PHJ_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    PHJ_CUDA(cudaEventCreate(&D->ev_r));
    PHJ_CUDA(cudaEventCreate(&D->ev_local));
    for (auto& e : D->ev_c) PHJ_CUDA(cudaEventCreate(&e));
    for (auto& e : D->ev_t) PHJ_CUDA(cudaEventCreate(&e));
    return PHJ_OK;
}

// world x local partitions: the split digit space. The tables are probed out of L2, so few, large local
// partitions are fine; what the choice trades is the length of the runs the scatter stores over
// NVLink (tile / digits tuples) against the size of one table.
int dist_plan(phj_dist* D) {
    const int W = D->world;
    if (W < 1 || W > kMaxRanks) return fail(PHJ_ERR_INVALID, "the number of GPUs must be in [1, %d]", kMaxRanks);
    // Every GPU owns the same number of split digits. For 2, 4, 8, 16 GPUs the digit is a bit field of the hash
    // (64 digits by default); for any other count it is hash % digits with digits = GPUs x floor(64 / GPUs)
    // (63, 60, 60, 63 for 3, 5, 6, 7 GPUs), the kernels' `%` path.
    uint32_t ndig = (64u / (uint32_t)W) * (uint32_t)W;
    if (D->cfg.partitions) {
        const uint64_t P = D->cfg.partitions;
        const bool pow2_world = (W & (W - 1)) == 0;
        if (P < (uint64_t)W || P > (uint64_t)kMaxSplitDigits || P % (uint64_t)W || (pow2_world && (P & (P - 1))))
            return fail(PHJ_ERR_INVALID, "sharded join: partitions = GPUs x local partitions must be a multiple of the "
                                         "%d GPUs in [%d, %d]%s (0 = choose)", W, W, kMaxSplitDigits,
                        pow2_world ? " and a power of two" : "");
        ndig = (uint32_t)P;
    }
    ndig = std::max<uint32_t>(ndig, (uint32_t)W);
    D->ndig = ndig;
    D->d_local = ndig / (uint32_t)W;
    D->b_local = (uint32_t)ilog2_ceil(D->d_local);
    D->K = D->cfg.split_chunks ? D->cfg.split_chunks : 4;
    if (D->K > (uint32_t)kMaxSplitChunks) return fail(PHJ_ERR_INVALID, "split_chunks must be <= %d", kMaxSplitChunks);
    // counting the later chunks behind the exchange is opt-in: a wash at 8 GPUs, slower at 2 (see the join)
    D->piecewise = (D->cfg.flags & PHJ_FLAG_COUNT_PIECEWISE) && !(D->cfg.flags & PHJ_FLAG_COUNT_UPFRONT);
    return PHJ_OK;
}

int dist_make_split(phj_dist* D) {
    phj_config sc = D->cfg;
    sc.algo = PHJ_ALGO_SHARD_SPLIT;
    sc.partitions = D->ndig;
    sc.radix_bits[0] = sc.radix_bits[1] = 0;
    sc.shard_shift = 0;
    sc.split_chunks = D->K;
    sc.split_ctas = 0;
    sc.num_gpus = 0;
    sc.flags = (D->cfg.flags & PHJ_FLAG_NO_TMA_STORE) | PHJ_FLAG_SPLIT_REMOTE_ONLY;
    sc.device = D->device;
    return phj_create(&sc, &D->split);
}

void dist_free(phj_dist* D) {
    if (!D) return;
    cudaSetDevice(D->device);
    if (D->split && D->split->stream) cudaStreamSynchronize(D->split->stream);
    if (D->sb) cudaStreamSynchronize(D->sb);
    if (D->sc) cudaStreamSynchronize(D->sc);
    for (int w = 0; w < 2; ++w)
        for (int r = 0; r < D->world; ++r)
            if (D->peer[w][r] && r != D->rank && !D->group) cudaIpcCloseMemHandle(D->peer[w][r]);
    if (D->comm && D->own_comm && nccl_api()) nccl_api()->CommDestroy(D->comm);
    for (int w = 0; w < 2; ++w)
        if (D->win[w]) cudaFree(D->win[w]);
    void* dev[] = {D->d_all_starts, D->d_peer, D->d_outd, D->d_lb, D->d_flags, D->d_xchg, D->d_pt, D->d_ptflags, D->d_pulls};
    for (void* p : dev)
        if (p) cudaFree(p);
    if (D->h_all_starts) cudaFreeHost(D->h_all_starts);
    if (D->h_flags) cudaFreeHost(D->h_flags);
    if (D->h_xchg) cudaFreeHost(D->h_xchg);
    if (D->sb) cudaStreamDestroy(D->sb);
    if (D->sc) cudaStreamDestroy(D->sc);
    for (auto& e : D->ev_cnt)
        if (e) cudaEventDestroy(e);
    if (D->ev_r) cudaEventDestroy(D->ev_r);
    if (D->ev_local) cudaEventDestroy(D->ev_local);
    for (auto& e : D->ev_c)
        if (e) cudaEventDestroy(e);
    for (auto& e : D->ev_t)
        if (e) cudaEventDestroy(e);
    if (D->split) phj_destroy(D->split);
    delete D;
}

int dist_set_relations(phj_dist* D, const void* build, size_t n_build, const void* probe, size_t n_probe,
                       bool device_resident) {
    PHJ_CUDA(cudaSetDevice(D->device));
    int rc = device_resident ? phj_bind_device(D->split, build, n_build, probe, n_probe)
                             : phj_upload(D->split, reinterpret_cast<const phj_tuple*>(build), n_build,
                                          reinterpret_cast<const phj_tuple*>(probe), n_probe);
    if (rc != PHJ_OK) return rc;
    D->n[0] = n_build;
    D->n[1] = n_probe;
    D->have_data = true;
    D->sized = false;  // new relations: count on the host once, keep windows that are large enough
    return PHJ_OK;
}

// ---- all ranks in one process -----------------------------------------------------------------------
void worker_main(phj_group::Worker* w) {
    std::unique_lock<std::mutex> lk(w->m);
    for (;;) {
        w->cv.wait(lk, [&] { return w->has_job || w->stop; });
        if (w->stop) return;
        std::function<int()> job = std::move(w->job);
        w->has_job = false;
        lk.unlock();
        g_error.clear();
        const int rc = job();
        lk.lock();
        w->rc = rc;
        w->error = g_error;
        w->done = true;
        w->cv.notify_all();
    }
}

// Run f(rank) on every rank's own host thread (the ranks synchronise with each other inside) and wait.
int group_run(phj_group* g, const std::function<int(int)>& f) {
    for (int r = 0; r < g->world; ++r) {
        phj_group::Worker* w = g->workers[r];
        std::lock_guard<std::mutex> lk(w->m);
        w->job = [&f, r] { return f(r); };
        w->has_job = true;
        w->done = false;
        w->cv.notify_all();
    }
    int rc = PHJ_OK;
    std::string first, all;
    for (int r = 0; r < g->world; ++r) {
        phj_group::Worker* w = g->workers[r];
        std::unique_lock<std::mutex> lk(w->m);
        w->cv.wait(lk, [&] { return w->done; });
        if (w->rc == PHJ_OK) continue;
        if (rc == PHJ_OK) {
            rc = w->rc;
            first = w->error;
        } else if (w->error == first || all.size() > 600) {
            continue;  // the ranks that only learnt of another rank's failure all say the same
        }
        all += (all.empty() ? "" : "; ") + ("GPU " + std::to_string(r) + ": " + w->error);
    }
    if (rc != PHJ_OK) g_error = all;
    return rc;
}

}  // namespace

void group_destroy(phj_group* g) {
    if (!g) return;
    for (auto* w : g->workers) {
        {
            std::lock_guard<std::mutex> lk(w->m);
            w->stop = true;
            w->cv.notify_all();
        }
        if (w->thread.joinable()) w->thread.join();
        delete w;
    }
    for (phj_dist* D : g->ranks) dist_free(D);
    for (phj_handle* h : g->replicas) phj_destroy(h);
    if (!g->comms.empty() && nccl_api())
        for (ncclComm_t c : g->comms)
            if (c) nccl_api()->CommDestroy(c);
    if (g->bar_ok) pthread_barrier_destroy(&g->bar);
    delete g;
}

int group_create(const phj_config* cfg, phj_group** out) {
    *out = nullptr;
    if (cfg->algo != PHJ_ALGO_RADIX_PARTITIONING && cfg->algo != PHJ_ALGO_NO_PARTITIONING)
        return fail(PHJ_ERR_INVALID, "num_gpus > 1 serves the radix-partitioning and the no-partitioning join");
    const int W = cfg->num_gpus;
    int ndev = 0;
    PHJ_CUDA(cudaGetDeviceCount(&ndev));
    if (W > kMaxRanks) return fail(PHJ_ERR_INVALID, "the number of GPUs must be in [1, %d]", kMaxRanks);
    if (W > ndev) return fail(PHJ_ERR_INVALID, "num_gpus = %d but only %d CUDA devices are visible", W, ndev);
    if (cfg->device < 0 || cfg->device + W > ndev)
        return fail(PHJ_ERR_INVALID, "devices %d .. %d are not all visible", cfg->device, cfg->device + W - 1);
    phj_group* g = new phj_group;
    g->world = W;
    auto cleanup = [&](int code) {
        group_destroy(g);
        return code;
    };
    if (cfg->algo == PHJ_ALGO_NO_PARTITIONING) {  // replicas of the build side: one plain handle per GPU
        for (int r = 0; r < W; ++r) {
            phj_config c = *cfg;
            c.num_gpus = 0;
            c.device = cfg->device + r;
            phj_handle* h = nullptr;
            const int rc = phj_create(&c, &h);
            if (rc != PHJ_OK) return cleanup(rc);
            g->replicas.push_back(h);
        }
        for (int r = 0; r < W; ++r) {
            auto* w = new phj_group::Worker;
            g->workers.push_back(w);
            w->thread = std::thread(worker_main, w);
        }
        *out = g;
        return PHJ_OK;
    }
    NcclApi* nc = nccl_api();
    if (!nc) return cleanup(fail(PHJ_ERR_CUDA, "num_gpus > 1 needs NCCL"));
    if (pthread_barrier_init(&g->bar, nullptr, (unsigned)W) != 0) return cleanup(fail(PHJ_ERR_NOMEM, "pthread_barrier_init failed"));
    g->bar_ok = true;
    int devs[kMaxRanks];
    for (int r = 0; r < W; ++r) devs[r] = cfg->device + r;
    if (cfg->device + W > ndev) return cleanup(fail(PHJ_ERR_INVALID, "devices %d .. %d are not all visible", cfg->device, cfg->device + W - 1));
    for (int r = 0; r < W; ++r) {  // peer access both ways (NCCL may already have enabled it)
        cudaSetDevice(devs[r]);
        for (int q = 0; q < W; ++q)
            if (q != r) {
                int can = 0;
                cudaDeviceCanAccessPeer(&can, devs[r], devs[q]);
                if (!can) return cleanup(fail(PHJ_ERR_CUDA, "GPU %d cannot access GPU %d's memory", devs[r], devs[q]));
                cudaError_t e = cudaDeviceEnablePeerAccess(devs[q], 0);
                if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled)
                    return cleanup(fail(PHJ_ERR_CUDA, "cudaDeviceEnablePeerAccess failed: %s", cudaGetErrorString(e)));
                cudaGetLastError();
            }
    }
    g->comms.assign(W, nullptr);
    {
        ncclResult_t r_ = nc->CommInitAll(g->comms.data(), W, devs);
        if (r_ != ncclSuccess) return cleanup(fail(PHJ_ERR_CUDA, "ncclCommInitAll failed: %s", nc->GetErrorString(r_)));
    }
    for (int r = 0; r < W; ++r) {
        phj_dist* D = new phj_dist;
        g->ranks.push_back(D);
        D->rank = r;
        D->world = W;
        D->device = devs[r];
        D->cfg = *cfg;
        D->group = g;
        D->comm = g->comms[r];
        int rc = dist_plan(D);
        if (rc == PHJ_OK) rc = dist_make_split(D);
        if (rc == PHJ_OK) rc = dist_alloc(D);
        if (rc != PHJ_OK) return cleanup(rc);
    }
    for (int r = 0; r < W; ++r) {
        auto* w = new phj_group::Worker;
        g->workers.push_back(w);
        w->thread = std::thread(worker_main, w);
    }
    *out = g;
    return PHJ_OK;
}

// Row shards: GPU r gets rows [r n / W, (r + 1) n / W) of both relations.
int group_set_relations(phj_group* g, const void* build, size_t n_build, const void* probe, size_t n_probe,
                        bool device_resident) {
    if (device_resident)
        return fail(PHJ_ERR_INVALID, "num_gpus > 1: relations come from host memory (phj_upload / phj_join_host)");
    const int W = g->world;
    if (!g->replicas.empty())  // no-partitioning: the whole build relation on every GPU, a row shard of the probe side
        return group_run(g, [&](int r) {
            const size_t p0 = n_probe * r / W, p1 = n_probe * (r + 1) / W;
            return phj_upload(g->replicas[r], reinterpret_cast<const phj_tuple*>(build), n_build,
                              reinterpret_cast<const phj_tuple*>(probe) + p0, p1 - p0);
        });
    return group_run(g, [&](int r) {
        const size_t b0 = n_build * r / W, b1 = n_build * (r + 1) / W;
        const size_t p0 = n_probe * r / W, p1 = n_probe * (r + 1) / W;
        return dist_set_relations(g->ranks[r], reinterpret_cast<const phj_tuple*>(build) + b0, b1 - b0,
                                  reinterpret_cast<const phj_tuple*>(probe) + p0, p1 - p0, false);
    });
}

int group_join(phj_group* g, phj_result* out) {
    std::vector<phj_result> res(g->world);
    const bool replicas = !g->replicas.empty();
    int rc = group_run(g, [&](int r) {
        return replicas ? phj_join(g->replicas[r], &res[r]) : dist_join_rank(g->ranks[r], &res[r]);
    });
    if (rc != PHJ_OK) return rc;
    *out = res[0];
    if (replicas) {  // every GPU counted its probe shard against the whole build relation
        out->gpus = (uint32_t)g->world;
        for (int r = 1; r < g->world; ++r) {
            out->matches += res[r].matches;
            out->build_ns = std::max(out->build_ns, res[r].build_ns);
            out->probe_ns = std::max(out->probe_ns, res[r].probe_ns);
        }
    }
    for (int r = 1; r < g->world; ++r) {  // the slowest rank defines the join's times
        out->total_ns = std::max(out->total_ns, res[r].total_ns);
        out->partition_ns = std::max(out->partition_ns, res[r].partition_ns);
        out->shuffle_ns = std::max(out->shuffle_ns, res[r].shuffle_ns);
        out->count_ns = std::max(out->count_ns, res[r].count_ns);
        out->join_ns = std::max(out->join_ns, res[r].join_ns);
        out->hbm_bytes_alg += res[r].hbm_bytes_alg;
        out->kernel_launches += res[r].kernel_launches;
        out->shuffle_bytes += res[r].shuffle_bytes;
        out->d2h_bytes += res[r].d2h_bytes;
    }
    return PHJ_OK;
}

// =================================================================================================
// C ABI of the sharded join, one process per GPU
// =================================================================================================
extern "C" {

int phj_nccl_unique_id(unsigned char* id128) {
    if (!id128) return fail(PHJ_ERR_INVALID, "id is null");
    NcclApi* nc = nccl_api();
    if (!nc) return fail(PHJ_ERR_CUDA, "NCCL is not available");
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId size");
    ncclUniqueId id;
    PHJ_NCCL(nc->GetUniqueId(&id));
    memcpy(id128, &id, 128);
    return PHJ_OK;
}

int phj_dist_create(const phj_config* config, int32_t rank, int32_t world, const unsigned char* id128,
                    phj_dist** out) {
    if (!out || !config || !id128) return fail(PHJ_ERR_INVALID, "null argument");
    *out = nullptr;
    if (config->algo != PHJ_ALGO_RADIX_PARTITIONING)
        return fail(PHJ_ERR_INVALID, "the sharded join is the radix-partitioning join");
    if (rank < 0 || rank >= world) return fail(PHJ_ERR_INVALID, "rank %d out of range [0, %d)", rank, world);
    NcclApi* nc = nccl_api();
    if (!nc) return fail(PHJ_ERR_CUDA, "the sharded join needs NCCL (libnccl.so.2)");
    phj_dist* D = new phj_dist;
    D->rank = rank;
    D->world = world;
    D->device = config->device;
    D->cfg = *config;
    auto cleanup = [&](int code) {
        dist_free(D);
        return code;
    };
    int rc = dist_plan(D);
    if (rc != PHJ_OK) return cleanup(rc);
    if (cudaSetDevice(D->device) != cudaSuccess) return cleanup(fail(PHJ_ERR_CUDA, "cudaSetDevice(%d) failed", D->device));
    ncclUniqueId id;
    memcpy(&id, id128, 128);
    {
        ncclResult_t r_ = nc->CommInitRank(&D->comm, world, id, rank);
        if (r_ != ncclSuccess) return cleanup(fail(PHJ_ERR_CUDA, "ncclCommInitRank failed: %s", nc->GetErrorString(r_)));
        D->own_comm = true;
    }
    if ((rc = dist_make_split(D)) != PHJ_OK) return cleanup(rc);
    if ((rc = dist_alloc(D)) != PHJ_OK) return cleanup(rc);
    *out = D;
    return PHJ_OK;
}

void phj_dist_destroy(phj_dist* d) {
    if (d && !d->group) dist_free(d);
}

int phj_dist_upload(phj_dist* d, const phj_tuple* build, size_t n_build, const phj_tuple* probe, size_t n_probe) {
    if (!d) return fail(PHJ_ERR_INVALID, "handle is null");
    return dist_set_relations(d, build, n_build, probe, n_probe, false);
}

int phj_dist_bind_device(phj_dist* d, const void* d_build, size_t n_build, const void* d_probe, size_t n_probe) {
    if (!d) return fail(PHJ_ERR_INVALID, "handle is null");
    return dist_set_relations(d, d_build, n_build, d_probe, n_probe, true);
}

int phj_dist_join(phj_dist* d, phj_result* out) {
    if (!d || !out) return fail(PHJ_ERR_INVALID, "handle or result is null");
    return dist_join_rank(d, out);
}

int phj_dist_kernel_times(phj_dist* d, const char** names, uint64_t* ns, uint32_t cap) {
    return d ? phj_kernel_times(d->split, names, ns, cap) : 0;
}

int phj_dist_kernel_timing(phj_dist* d, const char* filter) {
    if (!d) return fail(PHJ_ERR_INVALID, "handle is null");
    return phj_kernel_timing(d->split, filter);
}

int phj_dist_kernel_trace(phj_dist* d, const char** names, uint64_t* begin_ns, uint64_t* end_ns, uint32_t cap) {
    if (!d) return 0;
    phj_handle* h = d->split;
    uint32_t n = 0;
    for (int i = 0; i < h->n_ktimes && n < cap; ++i) {
        if (!h->ktimes[i].used) continue;
        names[n] = h->ktimes[i].name;
        begin_ns[n] = (uint64_t)(ev_ms(d->ev_t[0], h->ktimes[i].begin) * 1e6);
        end_ns[n] = (uint64_t)(ev_ms(d->ev_t[0], h->ktimes[i].end) * 1e6);
        ++n;
    }
    return (int)n;
}

int phj_dist_measure_peer_copy(phj_dist* d, uint64_t bytes, uint32_t repeats, uint64_t* ns_per_copy) {
    if (!d || !ns_per_copy) return fail(PHJ_ERR_INVALID, "null argument");
    if (!d->sized) return fail(PHJ_ERR_STATE, "the windows exist after the first join");
    PHJ_CUDA(cudaSetDevice(d->device));
    *ns_per_copy = 0;
    if (d->world < 2) return PHJ_OK;
    const int next = (d->rank + 1) % d->world;
    bytes = std::min<uint64_t>(bytes, 16 * std::min(d->caps[1][d->rank], d->caps[1][next]));
    cudaStream_t sa = d->split->stream;
    int rc = dist_host_barrier(d);  // everybody copies at the same time: every GPU sends and receives
    if (rc != PHJ_OK) return rc;
    PHJ_CUDA(cudaEventRecord(d->ev_t[6], sa));
    for (uint32_t i = 0; i < std::max<uint32_t>(repeats, 1); ++i)
        PHJ_CUDA(cudaMemcpyAsync(d->peer[1][next], d->win[1], bytes, cudaMemcpyDeviceToDevice, sa));
    PHJ_CUDA(cudaEventRecord(d->ev_t[7], sa));
    PHJ_CUDA(cudaStreamSynchronize(sa));
    *ns_per_copy = (uint64_t)(ev_ms(d->ev_t[6], d->ev_t[7]) * 1e6 / std::max<uint32_t>(repeats, 1));
    return dist_host_barrier(d);
}

int phj_dist_info(phj_dist* d, phj_dist_layout* out) {
    if (!d || !out) return fail(PHJ_ERR_INVALID, "null argument");
    memset(out, 0, sizeof(*out));
    out->world = (uint32_t)d->world;
    out->rank = (uint32_t)d->rank;
    out->digits = d->ndig;
    out->local_partitions = d->d_local;
    out->chunks = d->K;
    out->region_buckets = d->region_buckets;
    out->window_tuples[0] = d->caps[0][d->rank];
    out->window_tuples[1] = d->caps[1][d->rank];
    out->resizes = d->resizes;
    out->sent_remote_bytes = d->sent_remote_bytes;
    out->partitions_here = d->np;
    out->hot_count = d->n_hot;
    for (uint32_t i = 0; i < d->n_hot; ++i) out->hot_digits[i] = d->hot[i];
    return PHJ_OK;
}

int phj_dist_read_window(phj_dist* d, int32_t which, phj_tuple* out, uint64_t cap_tuples, uint64_t* bounds) {
    if (!d) return fail(PHJ_ERR_INVALID, "handle is null");
    if (which < 0 || which > 1) return fail(PHJ_ERR_INVALID, "which must be 0 (build) or 1 (probe)");
    if (!d->sized) return fail(PHJ_ERR_STATE, "no join has run yet");
    PHJ_CUDA(cudaSetDevice(d->device));
    const uint32_t np = d->np;
    const size_t nb = which == 0 ? np + 1 : (size_t)d->K * (np + 1);
    std::vector<uint64_t> b(nb);
    PHJ_CUDA(cudaMemcpy(b.data(), d->d_lb + (which == 0 ? 0 : np + 1), nb * 8, cudaMemcpyDeviceToHost));
    if (bounds) memcpy(bounds, b.data(), nb * 8);
    const uint64_t used = b[nb - 1];
    if (out) {
        if (used > cap_tuples) return fail(PHJ_ERR_INVALID, "the window holds %llu tuples, room for %llu", (unsigned long long)used, (unsigned long long)cap_tuples);
        if (used) PHJ_CUDA(cudaMemcpy(out, d->win[which], used * 16, cudaMemcpyDeviceToHost));
    }
    return PHJ_OK;
}

}  // extern "C"
