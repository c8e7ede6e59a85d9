// phj_kernels.cuh -- sm_100a kernels of the hash-join hot path.
//
// Reference loops these kernels replace (file:line under /root/reference/):
//   radix_histogram   scanTable                 src/RadixCluster/HashJoin.hpp:343-357
//   scan_lookback     createPrefixSumTable      src/RadixCluster/HashJoin.hpp:363-390
//                     + ComputePartitionsBoundaries                         :18-25
//   radix_scatter     partitionTable            src/RadixCluster/HashJoin.hpp:394-412
//   join_partitions   Join lambda               src/RadixCluster/HashJoin.hpp:258-323
//                     (+ LinearProbingHashTable::Insert/Get, src/HashTables/LinearProbing.hpp:114-180)
//   gt_build/gt_probe NoPartitioning Build/Probe src/NoPartitioning/HashJoin.hpp:76-126,128-187
//
// Vocabulary: a *segment* is what the reference calls a worker's batch: a contiguous slice of one
// relation (inside one parent partition for pass 2) that one CTA histograms and later scatters
// with private write cursors. `counts` is the reference's PrefixSumTable, laid out
// [parent][digit][segment] so that ONE flat exclusive scan yields every cursor and, at segment 0
// of each digit, the partition boundaries.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <type_traits>

#include "phj_hash.cuh"

#ifndef PHJ_SCAT_MINB
#define PHJ_SCAT_MINB 1
#endif
#ifndef PHJ_SCAT_PSCAN
#define PHJ_SCAT_PSCAN 1  // all-warp counter scan in radix_scatter
#endif
#ifndef PHJ_SCAT_ALLWRITE
#define PHJ_SCAT_ALLWRITE 1  // every lane of a digit group stores the warp counter
#endif
#ifndef PHJ_SCAT_PEERS4
#define PHJ_SCAT_PEERS4 1  // warp_peers in three ALU instructions per digit bit (inline PTX; 0: the C++ form, six)
#endif


namespace phj {

constexpr uint64_t kEmptyKey = 0x8000000000000000ULL;  // INT64_MIN marks a free table slot
constexpr uint32_t kFuse2MaxD2 = 64;  // fused pass-2 histogram: pass 2 of at most 6 bits (with a 6-bit pass 1)
constexpr uint32_t kFullD1 = 64, kFullD2 = 64;  // radix_histogram_full: two passes of at most 6 bits
constexpr uint32_t kMaxPieces = 18;  // sharded join: the build relation + up to 16 probe chunks

struct __align__(16) Segment {
    uint64_t begin, end;  // tuple range inside the relation's input array
    uint32_t cnt_index;   // counter index of (this segment, digit 0)
    uint32_t cnt_stride;  // counter stride between digits = #segments of the parent partition
    uint32_t rel;         // 0 build (R), 1 probe (S)
    uint32_t parent_first;  // bit 31: first segment of its parent; bits 0..30: parent partition id
};

// How a 64-bit hash becomes this pass's digit. full = POW2 ? h & pmask : h % modulus is the
// reference's partition id (src/Common/XXHasher.hpp:21); the pass takes bits [shift, shift+bits).
struct DigitFn {
    uint64_t pmask;
    uint64_t modulus;
    uint32_t shift;
    uint32_t mask;
};

template <bool POW2>
__device__ __forceinline__ uint32_t digit_of(uint64_t h, const DigitFn& f) {
    uint64_t full = POW2 ? (h & f.pmask) : (h % f.modulus);
    return (uint32_t)(full >> f.shift) & f.mask;
}

struct PassParams {
    const ulonglong2* in[2];
    ulonglong2* out[2];
    const Segment* segs;
    const uint32_t* nsegs;     // device-resident segment count (CTAs beyond it exit)
    uint32_t seg_first, seg_count;  // seg_count != 0: only segments [seg_first, seg_first + seg_count)
                               // (one row chunk of the multi-GPU split); else all *nsegs of them
    uint32_t* counts;          // histogram out (radix_histogram)
    const uint64_t* cursors;   // exclusive scan of counts (radix_scatter)
    uint64_t cursor_bias[2];   // flat scan spans R then S: S cursors are offset by |R|
    ulonglong2* const* outd[2];  // optional (shard split): per-digit output base, device array of
                               // ndigits pointers per relation; digit d's run goes to outd[d] + cursor
    uint64_t* bounds[2];       // partition boundaries of this pass's output, per relation
    uint32_t bounds_stride;    // digits per parent in `bounds` indexing
    uint32_t ndigits;          // digits this pass really has (<= the kernel's 1 << BITS)
    HashParams hp;
    DigitFn df;
    // Fused pass-2 histogram (radix_scatter<..., FUSE2 = true>, pass 1 only): while a tile is
    // staged, every tuple also counts into the pass-2 counter of the pass-2 segment its OUTPUT
    // position falls into, so pass 2 needs no histogram read of its own.
    const struct Parent2* parents2[2];  // per relation, one entry per pass-1 partition
    uint32_t* counts2;                  // pass-2 counters (zeroed before the launch)
    uint64_t seg_len2[2];               // pass-2 segment length per relation
    DigitFn df2;                        // pass-2 digit of the same hash
    uint32_t d2;                        // pass-2 digits
    // radix_histogram_full: per pass-1 segment, the counts of all (pass-1 digit, pass-2 digit) pairs
    uint32_t* hist12;                   // [segment][kFullD1 * kFullD2]
};

// Where pass-1 partition `parent` of a relation sits in pass 2's bookkeeping.
struct __align__(16) Parent2 {
    uint64_t lo;        // start of the partition in the pass-1 output (= its boundary)
    uint32_t cnt_base;  // index of counter (digit 0, segment 0) of this parent
    uint32_t nseg;      // pass-2 segments of this parent (= counter stride between digits)
};

// ---- small PTX helpers ---------------------------------------------------------------------------
// (tests/emu compiles this file for the host and brings host versions of exactly this block: every line of
// inline PTX of the kernels, except warp_peers' ballot below, is between this guard and its #endif)
#ifndef PHJ_PTX_HELPERS_PROVIDED
__device__ __forceinline__ ulonglong2 ld_stream_v2(const ulonglong2* p) {
    ulonglong2 v;
    asm volatile("ld.global.nc.L1::no_allocate.v2.u64 {%0, %1}, [%2];"
                 : "=l"(v.x), "=l"(v.y)
                 : "l"(p));
    return v;
}
// One 32-byte bucket (four keys) with ONE 256-bit load (LDG.E.256, sm_100): a divergent bucket read then
// costs the L1 one sector access instead of the two of a pair of 128-bit loads -- ncu showed the
// table probes bound by exactly that (l1tex throughput 96 %, DRAM and L2 half idle).
struct __align__(32) Bucket4 {
    unsigned long long k0, k1, k2, k3;
};
__device__ __forceinline__ Bucket4 ld_bucket4(const uint64_t* p) {
    Bucket4 b;
    asm volatile("ld.global.nc.v4.u64 {%0, %1, %2, %3}, [%4];"
                 : "=l"(b.k0), "=l"(b.k1), "=l"(b.k2), "=l"(b.k3)
                 : "l"(p));
    return b;
}
__device__ __forceinline__ uint64_t ld_stream_u64(const uint64_t* p) {
    uint64_t v;
    asm volatile("ld.global.nc.L1::no_allocate.u64 %0, [%1];" : "=l"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ void st_stream_v2(ulonglong2* p, const ulonglong2& v) {
    asm volatile("st.global.L1::no_allocate.v2.u64 [%0], {%1, %2};" ::"l"(p), "l"(v.x), "l"(v.y)
                 : "memory");
}
__device__ __forceinline__ uint32_t lanemask_lt() {
    uint32_t m;
    asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
    return m;
}
__device__ __forceinline__ uint64_t globaltimer_ns() {
    uint64_t t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
// Block barrier preceded by an explicit warp reconvergence. Measured on B200 (CUDA 12.9): after a
// region with per-lane loops (hash-table probing), ptxas may keep loop state in uniform registers
// and reach `bar.sync` without WARPSYNC; lanes of one warp then passed the barrier at different
// times and join_partitions lost matches. bar.warp.sync cannot be elided by the compiler.
__device__ __forceinline__ void cta_sync() {
    asm volatile("bar.warp.sync 0xffffffff;" ::: "memory");
    __syncthreads();
}
// TMA 1-D bulk store shared::cta -> global (UBLKCP in SASS). 16-byte aligned, size % 16 == 0.
__device__ __forceinline__ void bulk_store_s2g(void* gdst, const void* ssrc, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst),
                 "r"((uint32_t)__cvta_generic_to_shared(ssrc)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() {
    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}
// Full completion: the bulk stores' global writes are done (not merely their smem reads). A CTA must
// execute this before it exits -- measured on B200: with only the .read wait the next kernel on the
// stream could still observe the old contents of the destination.
__device__ __forceinline__ void bulk_wait_all0() {
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async_smem() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
#endif  // PHJ_PTX_HELPERS_PROVIDED

// Lanes of the warp whose digit equals this lane's. BALLOT = false: one match.any (a long-latency
// instruction on B200, tens of cycles of a per-SM unit); BALLOT = true: one vote per digit bit
// (NBITS votes + ~3 ALU each), which ncu shows to be much cheaper for <= 9 bits.
template <int NBITS, bool BALLOT>
__device__ __forceinline__ uint32_t warp_peers(uint32_t d) {
    if (!BALLOT) return __match_any_sync(0xffffffffu, d);
#if PHJ_SCAT_PEERS4
    // Spelled out in PTX so that ptxas sees six independent bit tests of one register: it then moves all of them into
    // predicates with ONE R2P and spends three ALU-pipe instructions per digit bit (VOTE, a predicated NOT for this
    // lane's polarity, an OR of `lanes that differ from me in this bit`). The C++ form below compiled to six per bit
    // (SHF + LOP3 + a second ISETP for the negated predicate + SEL + VOTE + LOP3), and the scatter's rank phase is
    // bound by exactly this pipe: measured 1275 / 1350 -> 1178 / 1228 us per pass at
    // 10 M x 200 M, partitions bit-identical (tools/ab_scatter.py peers4 check).
    uint32_t differ = 0;
#pragma unroll
    for (int b = 0; b < NBITS; ++b) {
        uint32_t m, pol;
        asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
            ".reg .b32 t;\n\t"
            "and.b32 t, %2, %3;\n\t"
            "setp.ne.u32 p, t, 0;\n\t"
            "vote.sync.ballot.b32 %0, p, 0xffffffff;\n\t"
            "selp.b32 %1, 0xffffffff, 0, p;\n\t"
            "}"
            : "=r"(m), "=r"(pol)
            : "r"(d), "r"(1u << b));
        differ |= m ^ pol;  // bit set: the lanes where it is clear (~m); bit clear: the lanes where it is set (m)
    }
    return ~differ;
#endif
    uint32_t peers = 0xffffffffu;
#pragma unroll
    for (int b = 0; b < NBITS; ++b) {
        const bool bit = (d & (1u << b)) != 0;             // LOP3 straight into a predicate
        const uint32_t m = __ballot_sync(0xffffffffu, bit);  // VOTE
        peers &= m ^ (bit ? 0u : 0xffffffffu);              // SEL + one 3-input LOP3
    }
    return peers;
}

// =================================================================================================
// K1  radix_histogram: one CTA per segment; per-warp private digit counters in shared memory,
// lanes of a warp that hit the same digit are merged with match.any so a Zipf heavy hitter costs
// one update per warp instead of 32 serialised ones, and no shared-memory atomics are needed.
// HBM: reads 16 B/tuple (the key's 32-byte sector is fetched whole either way).
// =================================================================================================
template <int BITS, int HASH, bool POW2, int TPB, int IPT, bool BALLOT>
__global__ void __launch_bounds__(TPB) radix_histogram(PassParams p) {
    constexpr int D = 1 << BITS;
    constexpr int NW = TPB / 32;
    constexpr int T = TPB * IPT;
    __shared__ uint32_t wc[NW][D + 1];  // [..][D] swallows the out-of-range lanes of a tail tile

    if (blockIdx.x >= (p.seg_count ? p.seg_count : *p.nsegs)) return;  // seg_count: one row chunk's segments
    const Segment seg = p.segs[p.seg_first + blockIdx.x];
    const ulonglong2* __restrict__ in = p.in[seg.rel];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t* wcw = wc[warp];

    for (int i = lane; i <= D; i += 32) wcw[i] = 0;
    __syncwarp();

    for (uint64_t tile = seg.begin; tile < seg.end; tile += T) {
        const uint64_t base = tile + (uint64_t)warp * (32 * IPT) + lane;
        uint64_t key[IPT];
#pragma unroll
        for (int i = 0; i < IPT; ++i) {
            const uint64_t idx = base + (uint64_t)i * 32;
            key[i] = idx < seg.end ? ld_stream_u64(reinterpret_cast<const uint64_t*>(in + idx)) : 0;
        }
#pragma unroll
        for (int i = 0; i < IPT; ++i) {
            const uint64_t idx = base + (uint64_t)i * 32;
            const uint32_t d = idx < seg.end ? digit_of<POW2>(hash_key<HASH>(key[i], p.hp), p.df) : D;
            const uint32_t peers = warp_peers<BITS + 1, BALLOT>(d);  // +1: the out-of-range digit D
            if (lane == __ffs(peers) - 1) wcw[d] += __popc(peers);
            __syncwarp();
        }
    }
    cta_sync();
    for (int d = threadIdx.x; d < (int)p.ndigits; d += TPB) {
        uint32_t sum = 0;
#pragma unroll
        for (int w = 0; w < NW; ++w) sum += wc[w][d];
        p.counts[seg.cnt_index + (uint64_t)d * seg.cnt_stride] = sum;
    }
}

// =================================================================================================
// K2  flat exclusive scan of the counters (uint32 counts -> uint64 cursors) in ONE launch (round 1 used two:
// per-chunk sums, then every CTA re-scanned the <= 1024 chunk sums for its carry-in and scanned its chunk).
// =================================================================================================
constexpr int kScanTpb = 1024;
constexpr int kScanIpt = 4;
constexpr int kScanChunk = kScanTpb * kScanIpt;

__device__ __forceinline__ uint64_t warp_incl_scan_u64(uint64_t v, int lane) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint64_t n = __shfl_up_sync(0xffffffffu, v, o);
        if (lane >= o) v += n;
    }
    return v;
}

// Block-wide exclusive scan over one value per thread (blockDim.x <= 1024). Returns the exclusive
// prefix; *total receives the block sum. `sh` needs 33 uint64.
__device__ __forceinline__ uint64_t block_excl_scan_u64(uint64_t v, uint64_t* sh, uint64_t* total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    uint64_t incl = warp_incl_scan_u64(v, lane);
    if (lane == 31) sh[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        uint64_t w = lane < nw ? sh[lane] : 0;
        uint64_t wi = warp_incl_scan_u64(w, lane);
        sh[lane] = wi - w;
        if (lane == 31) sh[32] = wi;
    }
    __syncthreads();
    uint64_t excl = incl - v + sh[warp];
    if (total) *total = sh[32];
    __syncthreads();
    return excl;
}

// scan_lookback: a CTA takes the next chunk (ticket), publishes
// the chunk's sum under this launch's epoch, adds up the sums of all earlier chunks as they appear (at most 1024,
// one per thread -- chunks are handed out in order, so every chunk a CTA waits for belongs to a CTA that is
// already running) and scans its own chunk. Saves a launch, a launch gap and a second read of the counters:
// 15 + 3 us -> 8 us per radix pass at 114 K counters (measured).
struct ScanState {
    unsigned long long sum[1024];  // chunk sums of the current launch
    uint32_t epoch[1024];          // launch that wrote sum[i]
    uint32_t ticket;               // next chunk; wraps to 0 with the last CTA of a launch
};

__global__ void __launch_bounds__(kScanTpb) scan_lookback(const uint32_t* __restrict__ counts,
                                                          const uint32_t* __restrict__ n_ptr,
                                                          uint64_t* __restrict__ cursors, ScanState* st,
                                                          uint32_t epoch) {
    __shared__ uint64_t sh[33];
    __shared__ uint64_t carry_sh;
    __shared__ uint32_t chunk_sh;
    if (threadIdx.x == 0) chunk_sh = atomicInc(&st->ticket, gridDim.x - 1);
    __syncthreads();
    const uint32_t chunk = chunk_sh;
    const uint32_t n = *n_ptr;
    const uint64_t base = (uint64_t)chunk * kScanChunk;
    // each thread owns kScanIpt consecutive counters
    const uint64_t first = base + (uint64_t)threadIdx.x * kScanIpt;
    uint32_t v[kScanIpt];
    uint64_t s = 0;
#pragma unroll
    for (int i = 0; i < kScanIpt; ++i) {
        v[i] = first + i < n ? counts[first + i] : 0;
        s += v[i];
    }
    uint64_t total;
    const uint64_t excl_in_chunk = block_excl_scan_u64(s, sh, &total);
    if (threadIdx.x == 0) {
        st->sum[chunk] = total;
        __threadfence();
        reinterpret_cast<volatile uint32_t*>(st->epoch)[chunk] = epoch;
    }
    if (base >= n) return;  // nothing to write (its sum, 0, is published for the chunks behind it)
    // carry-in: the sums of the chunks before this one
    uint64_t c = 0;
    {
        // warp-converged spin (see cta_sync()): every lane leaves the loop in the same trip
        const bool mine = threadIdx.x < chunk;
        bool ready = !mine;
        while (!__all_sync(0xffffffffu, ready))
            if (!ready) ready = reinterpret_cast<volatile uint32_t*>(st->epoch)[threadIdx.x] == epoch;
        __threadfence();
        if (mine) c = reinterpret_cast<volatile unsigned long long*>(st->sum)[threadIdx.x];
    }
    uint64_t carry;
    block_excl_scan_u64(c, sh, &carry);
    if (threadIdx.x == 0) carry_sh = carry;
    __syncthreads();
    uint64_t excl = excl_in_chunk + carry_sh;
#pragma unroll
    for (int i = 0; i < kScanIpt; ++i) {
        if (first + i < n) cursors[first + i] = excl;
        excl += v[i];
    }
}

// =================================================================================================
// K3  radix_scatter: one CTA per segment, tile by tile:
//   load T tuples into registers (warp-striped, 512 B per warp request)
//   rank   stable in-warp rank with match.any + per-warp digit counters (plain LDS/STS)
//   scan   per-digit scan across warps + across digits -> tile-local slot of every tuple
//   stage  tuples to their slot in shared memory (a counting sort of the tile)
//   flush  every digit's run goes out as one contiguous piece at the segment's private cursor:
//          TMA bulk store (cp.async.bulk.global.shared::cta, one per non-empty run) or
//          consecutive threads writing consecutive 16-byte tuples with st.global.v4
// Within a segment the order is preserved and segments are cut in input order, so the output
// equals the reference's stable partitionTable order (src/RadixCluster/HashJoin.hpp:400-407).
// HBM: reads 16 B/tuple, writes 16 B/tuple.
// =================================================================================================
template <int BITS, int TPB, int IPT>
struct ScatterSmem {
    static constexpr int D = 1 << BITS;
    static constexpr int NW = TPB / 32;
    static constexpr int T = TPB * IPT;
    static constexpr size_t stage_bytes = (size_t)T * 16;
    static constexpr size_t gcur_bytes = (size_t)D * 8;
    static constexpr size_t gbase_bytes = (size_t)D * 8;
    static constexpr size_t wc_bytes = (size_t)NW * (D + 1) * 4;
    static constexpr size_t dbase_bytes = (size_t)(D + 4) * 4;
    static constexpr size_t dtot_bytes = 64 * 4;  // PSCAN: per-digit tile totals (D <= 64), read in pairs
    static constexpr size_t sdig_bytes = (size_t)T * 2;
    static constexpr size_t total_tma =  // TMA bulk stores flush whole runs: no per-slot digit array
        (stage_bytes + gcur_bytes + gbase_bytes + wc_bytes + dbase_bytes + dtot_bytes + 16 * 4 + 64 + 15) / 16 * 16;
    static constexpr size_t total = total_tma + sdig_bytes;
    static constexpr size_t bytes(bool tma_store) { return tma_store ? total_tma : total; }
    // extra shared memory of the fused pass-2 histogram: counters + split positions + counter indices
    // 16-bit counters packed in pairs (flushed before they can overflow): with 32-bit counters the CTA
    // needed 127 KB and only ONE fit an SM, which is what made the first fused version slow
    static constexpr size_t fuse2_bytes = (size_t)2 * D * 64 * 2 + (size_t)D * 8 + (size_t)D * 8 + (size_t)D * 4;
};

template <int BITS, int HASH, bool POW2, int TPB, int IPT, bool TMA_STORE, bool BALLOT, bool FUSE2,
          int MINB = PHJ_SCAT_MINB>
__global__ void __launch_bounds__(TPB, MINB) radix_scatter(PassParams p) {
    using L = ScatterSmem<BITS, TPB, IPT>;
    constexpr int D = L::D, NW = L::NW, T = L::T;
    static_assert(D <= TPB, "one thread per digit in the scan step");
    extern __shared__ __align__(128) unsigned char smem_raw[];
    ulonglong2* stage = reinterpret_cast<ulonglong2*>(smem_raw);
    uint64_t* gcur = reinterpret_cast<uint64_t*>(smem_raw + L::stage_bytes);
    uint64_t* gbase = gcur + D;
    uint32_t* wtot_sh = reinterpret_cast<uint32_t*>(gbase + D);  // 16 entries: digit-scan warp totals
    uint32_t* dtot = wtot_sh + 16;                                // 64 entries, 8-byte aligned: PSCAN digit totals
    uint32_t* wc = dtot + 64;
    uint32_t* dbase = wc + NW * (D + 1);
    // PSCAN: a tile's counters are scanned by ALL warps (shuffles over the source-warp index, 32 / NW digits per
    // warp instruction) and every warp folds the digit bases into its own counter row: three CTA barriers per
    // tile instead of four, and no phase in which two warps work while the others wait.
    constexpr bool PSCAN = PHJ_SCAT_PSCAN && !FUSE2 && D <= 64 && (NW == 8 || NW == 16 || NW == 32);
    static_assert(T <= 0xffff, "tile slots fit 16 bits");
    uint16_t* sdig = reinterpret_cast<uint16_t*>(smem_raw + L::total_tma);  // only without TMA stores
    // FUSE2: [2][D][d2] pass-2 counters of this segment + per-digit split position / first segment
    uint32_t* h2 = reinterpret_cast<uint32_t*>(smem_raw + (TMA_STORE ? L::total_tma : L::total));
    uint64_t* split_pos = reinterpret_cast<uint64_t*>(h2 + D * kFuse2MaxD2);  // h2: 2 * D * 64 / 2 words
    uint32_t* h2_first = reinterpret_cast<uint32_t*>(split_pos + D);  // counter index of (digit 0, first seg)
    uint32_t* h2_stride = h2_first + D;
    // per tile: tile-local base of the digit | (tile slot at which its pass-2 segment changes) << 16
    uint32_t* dthr = h2_stride + D;
    static_assert(!FUSE2 || T <= 0x8000, "tile slots must fit 16 bits");

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    uint32_t* wcw = wc + warp * (D + 1);
    const uint32_t lt = lanemask_lt();
    const uint32_t nsegs = p.seg_count ? p.seg_count : *p.nsegs;
    if (PSCAN && tid < 64) dtot[tid] = 0;  // entries >= D are never written again

    // One CTA per segment by default; a smaller grid (the multi-GPU split, which leaves SMs to the
    // local join running beside it) makes every CTA walk several segments.
    for (uint32_t si = blockIdx.x; si < nsegs; si += gridDim.x) {
    if (si != blockIdx.x) cta_sync();  // the previous segment's last tile is staged and flushed
    const Segment seg = p.segs[p.seg_first + si];
    const ulonglong2* __restrict__ in = p.in[seg.rel];
    ulonglong2* __restrict__ out = p.out[seg.rel];

    // private write cursors of this segment (the reference's PrefixSumTable row + boundaries)
    if (tid < D) {
        uint64_t c = 0;
        if (tid < (int)p.ndigits) {
            c = p.cursors[seg.cnt_index + (uint64_t)tid * seg.cnt_stride] - p.cursor_bias[seg.rel];
            if (seg.parent_first >> 31)
                p.bounds[seg.rel][(uint64_t)(seg.parent_first & 0x7fffffffu) * p.bounds_stride + tid] = c;
        }
        gcur[tid] = c;
        if (FUSE2) {
            // Which pass-2 segment of parent `tid` this CTA's output run starts in, and where the
            // next one begins. The run is at most one pass-1 segment long (<= seg_len2), so it
            // touches at most two pass-2 segments: slot 0 and slot 1.
            Parent2 par{0, 0, 1};
            if (tid < (int)p.ndigits) par = p.parents2[seg.rel][tid];
            const uint64_t first = (c - par.lo) / p.seg_len2[seg.rel];
            split_pos[tid] = par.lo + (first + 1) * p.seg_len2[seg.rel];
            h2_first[tid] = par.cnt_base + (uint32_t)first;
            h2_stride[tid] = par.nseg;
        }
    }
    if (FUSE2) {
        for (int i = tid; i < D * (int)kFuse2MaxD2; i += TPB) h2[i] = 0;
    }

    // FUSE2: add the CTA's packed 16-bit pass-2 counts to the global counters and clear them.
    auto flush_h2 = [&]() {
        cta_sync();
        for (uint32_t w = tid; w < (uint32_t)D * kFuse2MaxD2; w += TPB) {
            const uint32_t pair = h2[w];
            if (pair) {
                h2[w] = 0;
#pragma unroll
                for (int half = 0; half < 2; ++half) {
                    const uint32_t c = (pair >> (16 * half)) & 0xffffu, i = 2 * w + half;
                    const uint32_t d2 = i % kFuse2MaxD2, d = (i / kFuse2MaxD2) % D, slot = i / (kFuse2MaxD2 * D);
                    if (c) atomicAdd(&p.counts2[h2_first[d] + d2 * h2_stride[d] + slot], c);
                }
            }
        }
        cta_sync();
    };
    uint32_t tiles_since_flush = 0;

    ulonglong2 v[IPT];
    auto load_tile = [&](uint64_t tile) {
        const ulonglong2* src = in + tile + (uint64_t)warp * (32 * IPT) + lane;
        if (tile + T <= seg.end) {
#pragma unroll
            for (int i = 0; i < IPT; ++i) v[i] = ld_stream_v2(src + i * 32);
        } else {
            const uint64_t base = tile + (uint64_t)warp * (32 * IPT) + lane;
#pragma unroll
            for (int i = 0; i < IPT; ++i)
                if (base + (uint64_t)i * 32 < seg.end) v[i] = ld_stream_v2(src + i * 32);
        }
    };
    if (seg.begin < seg.end) load_tile(seg.begin);

    for (uint64_t tile = seg.begin; tile < seg.end; tile += T) {
        const uint32_t n_valid = (uint32_t)min((uint64_t)T, seg.end - tile);
        const uint32_t n_mine = n_valid - min(n_valid, (uint32_t)(warp * (32 * IPT) + lane));  // > i*32 <=> valid
        if (FUSE2 && ++tiles_since_flush * T > 65535u) {  // a 16-bit counter could overflow in this tile
            flush_h2();
            tiles_since_flush = 1;
        }
        const bool full = n_valid == T;

        for (int i = lane; i <= D; i += 32) wcw[i] = 0;
        __syncwarp();

        // ---- rank (stable within the warp's 32*IPT consecutive tuples) ----
        // (Tried in round 2: first the IPT hash chains and votes as pure register work, then the counter updates as
        // one short LDS -> STS chain per round. Slower: 1.39 / 1.47 ms per pass against 1.29 / 1.36 ms -- the
        // interleaved rounds already overlap across the CTA's 16 warps, and the split costs registers.)
        uint32_t dr[IPT];  // digit (9 bits) | pass-2 digit << 9 (FUSE2, 7 bits) | rank << 16
        auto rank_round = [&](int i, auto is_full) {
            constexpr bool kFull = decltype(is_full)::value;
            const bool valid = kFull || (uint32_t)(i * 32) < n_mine;
            const uint64_t h = hash_key<HASH>(v[i].x, p.hp);
            const uint32_t d = valid ? digit_of<POW2>(h, p.df) : D;
            const uint32_t peers = warp_peers<kFull ? BITS : BITS + 1, BALLOT>(d);
            const uint32_t prev = wcw[d];
            __syncwarp();
#if PHJ_SCAT_ALLWRITE
            wcw[d] = prev + __popc(peers);  // every peer stores the same word: no leader election (BREV + FLO + ISETP)
#else
            if (lane == __ffs(peers) - 1) wcw[d] = prev + __popc(peers);
#endif
            __syncwarp();
            dr[i] = d | ((prev + __popc(peers & lt)) << 16);
            if (FUSE2) dr[i] |= digit_of<POW2>(h, p.df2) << 9;
        };
        if (full) {
#pragma unroll
            for (int i = 0; i < IPT; ++i) rank_round(i, std::true_type{});
        } else {
#pragma unroll
            for (int i = 0; i < IPT; ++i) rank_round(i, std::false_type{});
        }
        cta_sync();

        if constexpr (PSCAN) {
            // ---- scan across warps: lane = (digit slot q, source warp w_src) ----
            {
                const int w_src = lane % NW, q = lane / NW;
#pragma unroll
                for (int r = 0; r < (D + 31) / 32; ++r) {
                    const int d = r * 32 + warp + NW * q;  // shared-memory banks (w_src + d) % 32: no conflicts
                    const bool act = d < D && d < r * 32 + 32;
                    uint32_t* cell = wc + w_src * (D + 1) + (act ? d : 0);
                    const uint32_t c = act ? *cell : 0;
                    uint32_t incl = c;
#pragma unroll
                    for (int o = 1; o < NW; o <<= 1) {
                        const uint32_t n = __shfl_up_sync(0xffffffffu, incl, o);
                        if (w_src >= o) incl += n;
                    }
                    if (act) {
                        *cell = incl - c;  // tuples of digit d in the warps before w_src
                        if (w_src == NW - 1) dtot[d] = incl;
                    }
                }
            }
            if (TMA_STORE) bulk_wait_read0();  // previous tile's bulk stores have read `stage`
            cta_sync();
            // ---- scan across digits, redundantly in every warp (two digits per lane), folded into the warp's
            //      own counter row: wcw[d] = tile slot of this warp's first tuple of digit d ----
            {
                const uint2 t = reinterpret_cast<const uint2*>(dtot)[lane];
                const uint32_t s2 = t.x + t.y;
                uint32_t incl = s2;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const uint32_t n = __shfl_up_sync(0xffffffffu, incl, o);
                    if (lane >= o) incl += n;
                }
                const uint32_t ex0 = incl - s2, ex1 = ex0 + t.x;
                if (2 * lane < D) wcw[2 * lane] += ex0;
                if (2 * lane + 1 < D) wcw[2 * lane + 1] += ex1;
                if (warp < 2) {  // bookkeeping of the flush: warp 0 takes the even digits, warp 1 the odd ones
                    const int d = 2 * lane + warp;
                    if (d < D) {
                        const uint32_t ex = warp ? ex1 : ex0;
                        dbase[d] = ex;
                        if (d == D - 1) dbase[D] = n_valid;
                        const uint64_t g = gcur[d];
                        gbase[d] = g - ex;  // out index of tile slot j of this digit = gbase + j
                        gcur[d] = g + (warp ? t.y : t.x);
                    }
                }
                __syncwarp();
            }
        } else {
        // ---- per-digit scan across warps, then across digits (32-bit, two barriers) ----
        uint32_t tot = 0, incl = 0;
        if (tid < D) {
#pragma unroll
            for (int w = 0; w < NW; ++w) {
                const uint32_t c = wc[w * (D + 1) + tid];
                wc[w * (D + 1) + tid] = tot;
                tot += c;
            }
            incl = tot;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t n = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += n;
            }
            if (lane == 31) wtot_sh[warp] = incl;
        }
        if (TMA_STORE) bulk_wait_read0();  // previous tile's bulk stores have read `stage`
        cta_sync();
        if (tid < D) {
            uint32_t excl = incl - tot;
#pragma unroll
            for (int w = 0; w < (D + 31) / 32; ++w)
                if (w < warp) excl += wtot_sh[w];
            dbase[tid] = excl;
            if (tid == D - 1) dbase[D] = n_valid;
            gbase[tid] = gcur[tid] - excl;  // out index of tile slot j of this digit = gbase + j
            if (FUSE2) {
                // tile slot j of this digit goes to the run's second pass-2 segment iff gbase + j >=
                // split_pos, i.e. j >= thr: one packed 32-bit word per digit for the stage step
                // instead of two 64-bit loads and a 64-bit compare per tuple
                const int64_t t = (int64_t)(split_pos[tid] - gbase[tid]);
                const uint32_t thr = t < 0 ? 0u : t > 0xffff ? 0xffffu : (uint32_t)t;
                dthr[tid] = excl | (thr << 16);
            }
            gcur[tid] += tot;
        }
        cta_sync();
        }  // !PSCAN

        // ---- stage ----
#pragma unroll
        for (int i = 0; i < IPT; ++i) {
            const uint32_t d = dr[i] & 0x1ffu;
            if (full || d < D) {
                uint32_t db = 0;  // PSCAN: the digit's base is already inside wcw[d]
                if constexpr (!PSCAN) db = FUSE2 ? dthr[d] : dbase[d];
                const uint32_t pos = (FUSE2 ? db & 0xffffu : db) + wcw[d] + (dr[i] >> 16);
                stage[pos] = v[i];
                if (!TMA_STORE) sdig[pos] = (uint16_t)d;
                if constexpr (FUSE2) {
                    const uint32_t slot = pos >= (db >> 16);
                    const uint32_t c2 = (slot * D + d) * kFuse2MaxD2 + ((dr[i] >> 9) & 0x7fu);
                    atomicAdd(&h2[c2 >> 1], 1u << ((c2 & 1) * 16));
                }
            }
        }
        // prefetch the next tile while this one is flushed
        if (tile + T < seg.end) load_tile(tile + T);
        if (TMA_STORE) fence_proxy_async_smem();
        cta_sync();

        // ---- flush ----
        if (TMA_STORE) {
            if (tid < D) {
                const uint32_t b = dbase[tid], e = dbase[tid + 1];
                if (e > b) {
                    ulonglong2* od = p.outd[seg.rel] ? p.outd[seg.rel][tid] : out;  // maybe a peer GPU
                    bulk_store_s2g(od + gbase[tid] + b, stage + b, (e - b) * 16u);
                }
                bulk_commit();
            }
        } else {
            for (uint32_t j = tid; j < n_valid; j += TPB) {
                const uint32_t d = sdig[j];
                ulonglong2* od = p.outd[seg.rel] ? p.outd[seg.rel][d] : out;
                st_stream_v2(od + gbase[d] + j, stage[j]);
            }
        }
        // The next iteration's barriers separate this flush from the next stage step.
    }
    if (FUSE2) flush_h2();  // this segment's remaining pass-2 counts
    }  // segments of this CTA
    if (TMA_STORE) bulk_wait_all0();
}

// =================================================================================================
// K1b  radix_histogram_lanes (default histogram for passes of <= 6 bits).
//
// ncu on B200 showed the match.any of K1 to be its bottleneck (the instruction consuming its result
// carried ~45 % of the stall samples; 1.13 ms for a 3.36 GB read). This variant needs no
// cross-lane operation per tuple: every LANE owns private 16-bit digit counters in shared memory,
// laid out [digit pair][lane][2] so that a lane only ever touches its own bank (conflict-free).
// The lanes' counters are reduced once per segment. Measured: 0.485 ms = 6.9 TB/s.
// (The same idea was tried for the scatter -- rank = lane-private counter, then a per-digit scan
// across lanes -- and lost to the ballot ranking below: 2.2-2.9 ms vs 1.5 ms per pass.)
// =================================================================================================
template <int BITS>
__device__ __forceinline__ uint32_t lane_counter_offset(uint32_t d, int lane) {
    // byte offset of this lane's counter for digit d inside its warp's block
    return (((d >> 1) * 32 + lane) << 2) + ((d & 1) << 1);
}

template <int BITS, int TPB>
struct HistLanesSmem {
    static constexpr size_t total = ((size_t)(TPB / 32) * ((1 << BITS) / 2) * 32 + (1 << BITS)) * 4;
};

template <int BITS, int HASH, bool POW2, int TPB, int IPT>
__global__ void __launch_bounds__(TPB) radix_histogram_lanes(PassParams p) {
    constexpr int D = 1 << BITS;
    constexpr int NW = TPB / 32;
    constexpr int T = TPB * IPT;
    constexpr int WARP_WORDS = (D / 2) * 32;  // uint32 words per warp
    static_assert(BITS >= 1 && BITS <= 6, "lane-private counters are for fan-outs up to 64");
    extern __shared__ __align__(128) unsigned char smem_raw[];  // HistLanesSmem<BITS, TPB>::total bytes
    uint32_t* cnt = reinterpret_cast<uint32_t*>(smem_raw);      // [NW][WARP_WORDS]
    uint32_t* total = cnt + NW * WARP_WORDS;                    // [D]

    if (blockIdx.x >= (p.seg_count ? p.seg_count : *p.nsegs)) return;  // seg_count: one row chunk's segments
    const Segment seg = p.segs[p.seg_first + blockIdx.x];
    const ulonglong2* __restrict__ in = p.in[seg.rel];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned char* mine = reinterpret_cast<unsigned char*>(cnt + warp * WARP_WORDS);

    for (int i = lane; i < WARP_WORDS; i += 32) cnt[warp * WARP_WORDS + i] = 0;
    if (threadIdx.x < D) total[threadIdx.x] = 0;
    __syncwarp();

    auto count_key = [&](uint64_t key) {
        const uint32_t d = digit_of<POW2>(hash_key<HASH>(key, p.hp), p.df);
        uint16_t* c = reinterpret_cast<uint16_t*>(mine + lane_counter_offset<BITS>(d, lane));
        *c = (uint16_t)(*c + 1);
    };
    uint64_t tile = seg.begin;
    for (; tile + T <= seg.end; tile += T) {  // full tiles: no bounds checks
        const uint64_t* src = reinterpret_cast<const uint64_t*>(in + tile + (uint64_t)warp * (32 * IPT) + lane);
        uint64_t key[IPT];
#pragma unroll
        for (int i = 0; i < IPT; ++i) key[i] = ld_stream_u64(src + (size_t)i * 64);
#pragma unroll
        for (int i = 0; i < IPT; ++i) count_key(key[i]);
    }
    if (tile < seg.end) {  // tail tile
        const uint64_t base = tile + (uint64_t)warp * (32 * IPT) + lane;
#pragma unroll
        for (int i = 0; i < IPT; ++i) {
            const uint64_t idx = base + (uint64_t)i * 32;
            if (idx < seg.end) count_key(ld_stream_u64(reinterpret_cast<const uint64_t*>(in + idx)));
        }
    }
    __syncwarp();
    // reduce the 32 lanes' counters of every digit pair, then across warps
    for (int e = 0; e < D / 2; ++e) {
        const uint32_t w = cnt[warp * WARP_WORDS + e * 32 + lane];
        uint32_t lo = w & 0xffffu, hi = w >> 16;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            lo += __shfl_xor_sync(0xffffffffu, lo, o);
            hi += __shfl_xor_sync(0xffffffffu, hi, o);
        }
        if (lane == 0) {
            atomicAdd(&total[2 * e], lo);
            if (2 * e + 1 < D) atomicAdd(&total[2 * e + 1], hi);
        }
    }
    cta_sync();
    for (int d = threadIdx.x; d < (int)p.ndigits; d += TPB)
        p.counts[seg.cnt_index + (uint64_t)d * seg.cnt_stride] = total[d];
}

// =================================================================================================
// K1c  radix_histogram_lanes8: the lane-private idea for fan-outs of 128 / 256 (the multi-GPU split
// whose digit is owner rank x local pass-1 digit). Counters are 8 bits wide -- [digit / 4][lane]
// words, one byte per digit, a lane still only touches its own bank -- so a warp's block is
// D * 32 bytes (8 KB for 256 digits). A lane adds at most IPT per tile; every FLUSH_TILES tiles
// (IPT * FLUSH_TILES <= 255) the warp folds its bytes into the CTA's 32-bit totals and clears them.
// =================================================================================================
template <int BITS, int TPB>
struct HistLanes8Smem {
    static constexpr size_t total = ((size_t)(TPB / 32) * ((1 << BITS) / 4) * 32 + (1 << BITS)) * 4;
};

template <int BITS, int HASH, bool POW2, int TPB, int IPT>
__global__ void __launch_bounds__(TPB) radix_histogram_lanes8(PassParams p) {
    constexpr int D = 1 << BITS;
    constexpr int NW = TPB / 32;
    constexpr int T = TPB * IPT;
    constexpr int WARP_WORDS = (D / 4) * 32;
    constexpr int FLUSH_TILES = 255 / IPT;
    static_assert(BITS >= 2 && BITS <= 8, "byte counters: fan-outs up to 256");
    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint32_t* cnt = reinterpret_cast<uint32_t*>(smem_raw);  // [NW][WARP_WORDS]
    uint32_t* total = cnt + NW * WARP_WORDS;                // [D]

    if (blockIdx.x >= (p.seg_count ? p.seg_count : *p.nsegs)) return;  // seg_count: one row chunk's segments
    const Segment seg = p.segs[p.seg_first + blockIdx.x];
    const ulonglong2* __restrict__ in = p.in[seg.rel];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t* mine_w = cnt + warp * WARP_WORDS;
    unsigned char* mine = reinterpret_cast<unsigned char*>(mine_w);

    for (int i = lane; i < WARP_WORDS; i += 32) mine_w[i] = 0;
    for (int i = threadIdx.x; i < D; i += TPB) total[i] = 0;
    cta_sync();

    auto count_key = [&](uint64_t key) {
        const uint32_t d = digit_of<POW2>(hash_key<HASH>(key, p.hp), p.df);
        unsigned char* c = mine + ((((d >> 2) * 32 + lane) << 2) | (d & 3));
        *c = (unsigned char)(*c + 1);
    };
    auto flush = [&]() {
        __syncwarp();
        for (int e = 0; e < D / 4; ++e) {
            const uint32_t w = mine_w[e * 32 + lane];
            mine_w[e * 32 + lane] = 0;
            uint32_t even = w & 0x00ff00ffu, odd = (w >> 8) & 0x00ff00ffu;  // two 16-bit sums each
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                even += __shfl_xor_sync(0xffffffffu, even, o);
                odd += __shfl_xor_sync(0xffffffffu, odd, o);
            }
            if (lane < 4) {
                const uint32_t v = lane == 0 ? (even & 0xffffu) : lane == 1 ? (odd & 0xffffu)
                                 : lane == 2 ? (even >> 16) : (odd >> 16);
                if (v) atomicAdd(&total[4 * e + lane], v);
            }
        }
        __syncwarp();
    };
    uint64_t tile = seg.begin;
    int since_flush = 0;
    for (; tile + T <= seg.end; tile += T) {
        const uint64_t* src = reinterpret_cast<const uint64_t*>(in + tile + (uint64_t)warp * (32 * IPT) + lane);
        uint64_t key[IPT];
#pragma unroll
        for (int i = 0; i < IPT; ++i) key[i] = ld_stream_u64(src + (size_t)i * 64);
#pragma unroll
        for (int i = 0; i < IPT; ++i) count_key(key[i]);
        if (++since_flush == FLUSH_TILES) {
            flush();
            since_flush = 0;
        }
    }
    if (tile < seg.end) {
        const uint64_t base = tile + (uint64_t)warp * (32 * IPT) + lane;
#pragma unroll
        for (int i = 0; i < IPT; ++i) {
            const uint64_t idx = base + (uint64_t)i * 32;
            if (idx < seg.end) count_key(ld_stream_u64(reinterpret_cast<const uint64_t*>(in + idx)));
        }
    }
    flush();
    cta_sync();
    for (int d = threadIdx.x; d < (int)p.ndigits; d += TPB)
        p.counts[seg.cnt_index + (uint64_t)d * seg.cnt_stride] = total[d];
}

// =================================================================================================
// K1d  radix_histogram_full: ONE read of the input yields the histograms of BOTH passes (two passes
// of <= 6 bits). Every CTA counts its pass-1 segment into a [pass-1 digit][pass-2 digit] table of
// 4096 32-bit counters in shared memory (one shared-memory atomic per tuple: with 4096 bins two lanes
// rarely meet in a counter, unlike a 64-bin table) and writes (a) the pass-1 counts = row sums, as
// radix_histogram_lanes does, and (b) the table itself to hist12[segment]. The pass-1 scatter is
// stable and segments are input slices, so in the pass-1 output partition d1 is the concatenation
// over segments s of run (s, d1), whose pass-2 digit counts are exactly hist12[s][d1][*]: pass 2
// gets its counters by summing rows (pass2_align_counts) instead of reading its input again
// or having the pass-1 scatter count (FUSE2). The kernel stays HBM-bound: the atomics hide behind
// the loads. HBM: reads 16 B/tuple (8 used), writes 16 KB per segment.
// =================================================================================================
template <int HASH, int TPB, int IPT>
__global__ void __launch_bounds__(TPB) radix_histogram_full(PassParams p) {
    constexpr int T = TPB * IPT;
    constexpr int NB = kFullD1 * kFullD2;
    __shared__ uint32_t cnt[NB];

    if (blockIdx.x >= (p.seg_count ? p.seg_count : *p.nsegs)) return;  // seg_count: one row chunk's segments
    const Segment seg = p.segs[p.seg_first + blockIdx.x];
    const ulonglong2* __restrict__ in = p.in[seg.rel];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;

    for (int i = threadIdx.x; i < NB; i += TPB) cnt[i] = 0;
    cta_sync();

    auto count_key = [&](uint64_t key) {
        const uint64_t h = hash_key<HASH>(key, p.hp);
        atomicAdd(&cnt[digit_of<true>(h, p.df) * kFullD2 + digit_of<true>(h, p.df2)], 1u);
    };
    uint64_t tile = seg.begin;
    for (; tile + T <= seg.end; tile += T) {  // full tiles: no bounds checks
        const uint64_t* src = reinterpret_cast<const uint64_t*>(in + tile + (uint64_t)warp * (32 * IPT) + lane);
        uint64_t key[IPT];
#pragma unroll
        for (int i = 0; i < IPT; ++i) key[i] = ld_stream_u64(src + (size_t)i * 64);
#pragma unroll
        for (int i = 0; i < IPT; ++i) count_key(key[i]);
    }
    if (tile < seg.end) {  // tail tile
        const uint64_t base = tile + (uint64_t)warp * (32 * IPT) + lane;
#pragma unroll
        for (int i = 0; i < IPT; ++i) {
            const uint64_t idx = base + (uint64_t)i * 32;
            if (idx < seg.end) count_key(ld_stream_u64(reinterpret_cast<const uint64_t*>(in + idx)));
        }
    }
    cta_sync();
    // pass-1 counts: row sums (column index rotated by the row: the 64 threads hit 32 distinct banks twice)
    if (threadIdx.x < p.ndigits) {
        uint32_t sum = 0;
#pragma unroll 8
        for (uint32_t k = 0; k < kFullD2; ++k) sum += cnt[threadIdx.x * kFullD2 + ((k + threadIdx.x) & (kFullD2 - 1))];
        p.counts[seg.cnt_index + (uint64_t)threadIdx.x * seg.cnt_stride] = sum;
    }
    uint32_t* __restrict__ row = p.hist12 + (uint64_t)blockIdx.x * NB;
    for (int i = threadIdx.x; i < NB; i += TPB) row[i] = cnt[i];
}

// =================================================================================================
// plan_pass2: cut every pass-1 partition ("parent") of both relations into segments for pass 2.
// Single CTA; runs on the device so the pipeline never waits for the host.
// =================================================================================================
struct Plan2Params {
    // pass-1 boundaries come straight from the scanned cursors: boundary(rel, d) =
    // cursors[cnt_base1[rel] + d * nseg1[rel]] - bias[rel]; they are also written to bounds1.
    const uint64_t* cursors;
    uint32_t cnt_base1[2], nseg1[2];
    uint64_t bias[2];
    uint64_t* bounds1[2];  // out: D1 + 1 boundaries per relation (last = n, preset by the host)
    Parent2* parents2[2];  // out
    uint64_t seg_len[2];   // pass-2 segment length (host-chosen, a multiple of the tile)
    uint64_t n[2];
    uint32_t d1;          // parents per relation
    uint32_t d2;          // digits of pass 2
    uint32_t tile;        // tuples per tile of the pass-2 kernels
    uint32_t target_segs[2];
    Segment* segs;        // out
    uint32_t* nsegs;      // out: total segments
    uint32_t* ncounts;    // out: total counters = d2 * nsegs
    uint32_t max_segs;
    uint32_t* truncated;  // set when the parents need more than max_segs segments
    uint64_t* bounds2[2];  // pass-2 boundaries: those of EMPTY parents are written here (no segment ever does)
    unsigned long long* zero_words;  // the join's match counter and its oversize / truncated word: cleared here
};

__global__ void __launch_bounds__(1024) plan_pass2(Plan2Params p) {
    __shared__ uint64_t sh[33];
    __shared__ uint32_t rel_base_sh;
    uint32_t seg_base = 0;  // segments emitted by previous relations
    if (threadIdx.x == 0 && p.zero_words) p.zero_words[0] = p.zero_words[1] = 0;
    for (int rel = 0; rel < 2; ++rel) {
        const uint64_t n = p.n[rel];
        const uint64_t seg_len = p.seg_len[rel];
        uint64_t lo = 0, len = 0;
        uint32_t ns = 0;
        if (threadIdx.x < p.d1) {
            auto boundary = [&](uint32_t d) -> uint64_t {
                if (d >= p.d1 || p.nseg1[rel] == 0) return d >= p.d1 ? n : 0;
                return p.cursors[p.cnt_base1[rel] + (uint64_t)d * p.nseg1[rel]] - p.bias[rel];
            };
            lo = boundary(threadIdx.x);
            const uint64_t hi = boundary(threadIdx.x + 1);
            len = hi - lo;
            ns = (uint32_t)((len + seg_len - 1) / seg_len);
            p.bounds1[rel][threadIdx.x] = lo;
            if (len == 0 && p.bounds2[rel])  // no segment of this parent will write its boundaries
                for (uint32_t d = 0; d < p.d2; ++d) p.bounds2[rel][(uint64_t)threadIdx.x * p.d2 + d] = lo;
        }
        uint64_t total;
        const uint32_t first = (uint32_t)block_excl_scan_u64(ns, sh, &total);
        if (threadIdx.x < p.d1) {
            Parent2 par;
            par.lo = lo;
            par.cnt_base = (seg_base + first) * p.d2;
            par.nseg = ns;
            p.parents2[rel][threadIdx.x] = par;
            for (uint32_t s = 0; s < ns; ++s) {
                const uint32_t gi = seg_base + first + s;
                if (gi >= p.max_segs) break;
                Segment sg;
                sg.begin = lo + (uint64_t)s * seg_len;
                sg.end = min(lo + len, sg.begin + seg_len);
                sg.cnt_index = (seg_base + first) * p.d2 + s;
                sg.cnt_stride = ns;
                sg.rel = rel;
                sg.parent_first = threadIdx.x | (s == 0 ? 0x80000000u : 0u);
                p.segs[gi] = sg;
            }
        }
        if (threadIdx.x == 0) rel_base_sh = seg_base + (uint32_t)total;
        __syncthreads();
        seg_base = rel_base_sh;
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        *p.nsegs = min(seg_base, p.max_segs);
        *p.ncounts = min(seg_base, p.max_segs) * p.d2;
        if (seg_base > p.max_segs) *p.truncated = 1;
    }
}

// =================================================================================================
// With radix_histogram_full, pass 2's segments are aligned to RUNS of the pass-1 output: run (s, d)
// = the tuples of pass-1 digit d that pass-1 segment s wrote, starting at its scanned cursor.
// Segment j of parent d keeps plan_pass2's nominal window [lo + j L, lo + (j + 1) L) but holds exactly
// the runs that START inside it, so its pass-2 digit counts are sums of hist12 rows.
// pass2_align_counts moves every segment's begin / end to those run starts (binary search over the parent's
// cursors) and adds the rows up into the pass-2 counters (round 1: two launches, a memset and atomics).
// =================================================================================================
// 64 threads per pass-2 segment. All of them find the runs that
// start inside the segment's window ([a0, a1) by binary search over the parent's cursors, the loads are
// broadcasts), thread 0 writes the aligned segment, and thread d2 adds up hist12[s][d][d2] over those runs:
// every pass-2 counter is written exactly once, so no memset and no atomic. 11 + 33 us -> see DESIGN.md.
struct AlignCounts2Params {
    const uint64_t* cursors;  // scanned pass-1 cursors (as in Plan2Params)
    uint32_t cnt_base1[2], nseg1[2], seg_first1[2];
    uint64_t bias[2];
    const uint64_t* bounds1[2];  // d1 + 1 per relation
    const Parent2* parents2[2];
    uint64_t seg_len[2];
    Segment* segs;
    const uint32_t* nsegs;
    const uint32_t* hist12;  // [pass-1 segment][kFullD1][kFullD2]
    uint32_t d2;
    uint32_t* counts2;
};

__global__ void __launch_bounds__(256) pass2_align_counts(AlignCounts2Params p) {
    const uint32_t gi = blockIdx.x * (256 / kFullD2) + threadIdx.x / kFullD2, d2 = threadIdx.x % kFullD2;
    if (gi >= *p.nsegs) return;
    Segment sg = p.segs[gi];
    const uint32_t rel = sg.rel, d = sg.parent_first & 0x7fffffffu;
    const Parent2 par = p.parents2[rel][d];
    const uint32_t j = sg.cnt_index - par.cnt_base, n1 = p.nseg1[rel];
    const uint64_t lo = par.lo, hi = p.bounds1[rel][d + 1], bias = p.bias[rel];
    const uint64_t* __restrict__ cur = p.cursors + p.cnt_base1[rel] + (uint64_t)d * n1;
    // first run that starts at or after a target position; the two searches advance together
    const uint64_t t0 = lo + (uint64_t)j * p.seg_len[rel], t1 = t0 + p.seg_len[rel];
    uint32_t a0 = 0, b0 = n1, a1 = 0, b1 = n1;
    while (a0 < b0 || a1 < b1) {
        const uint32_t m0 = min((a0 + b0) >> 1, n1 - 1), m1 = min((a1 + b1) >> 1, n1 - 1);
        const uint64_t c0 = cur[m0] - bias, c1 = cur[m1] - bias;
        if (a0 < b0) {
            if (c0 >= t0) b0 = m0;
            else a0 = m0 + 1;
        }
        if (a1 < b1) {
            if (c1 >= t1) b1 = m1;
            else a1 = m1 + 1;
        }
    }
    const bool last = j + 1 == par.nseg;
    if (j == 0) a0 = 0;
    if (last) a1 = n1;
    if (d2 == 0) {
        sg.begin = j == 0 ? lo : a0 == n1 ? hi : cur[a0] - bias;
        sg.end = last ? hi : a1 == n1 ? hi : cur[a1] - bias;
        p.segs[gi] = sg;
    }
    if (d2 >= p.d2) return;
    const uint32_t* __restrict__ hrow =
        p.hist12 + ((uint64_t)p.seg_first1[rel] + a0) * (kFullD1 * kFullD2) + d * kFullD2 + d2;
    uint32_t acc0 = 0, acc1 = 0, acc2 = 0, acc3 = 0;
    uint32_t s = a0;
    for (; s + 4 <= a1; s += 4, hrow += 4 * kFullD1 * kFullD2) {
        acc0 += hrow[0];
        acc1 += hrow[kFullD1 * kFullD2];
        acc2 += hrow[2 * kFullD1 * kFullD2];
        acc3 += hrow[3 * kFullD1 * kFullD2];
    }
    for (; s < a1; ++s, hrow += kFullD1 * kFullD2) acc0 += hrow[0];
    p.counts2[par.cnt_base + d2 * par.nseg + j] = acc0 + acc1 + acc2 + acc3;
}

// =================================================================================================
// K4+K5  join_partitions: the probe relation's partitioned array is cut into gridDim.x equal
// slices; a CTA walks the partitions its slice overlaps, builds the partition's build-side keys
// into a shared-memory table of 32-byte buckets (four keys; 64-bit atomicCAS on the first free
// slot, overflow into the next bucket; duplicates collapse: the join counts probe tuples with
// >= 1 match) and streams its share of the probe partition through it, one bucket read per step.
// Equal slices make skew a non-issue: the partition holding a Zipf heavy hitter is simply probed
// by many CTAs, each rebuilding the (L2-resident) build partition. Partitions whose build side
// does not fit the table are left to gt_build/gt_probe and counted in *oversize.
// HBM: reads 16 B per probe tuple + ~16 B per build tuple.
// =================================================================================================
struct JoinParams {
    const ulonglong2* build;  // partitioned R
    const ulonglong2* probe;  // partitioned S
    const uint64_t* bounds_build;  // npart + 1
    const uint64_t* bounds_probe;
    uint64_t n_probe;
    uint64_t slice_len, slice_rem;  // n_probe / gridDim.x and n_probe % gridDim.x
    uint32_t npart;
    uint32_t bucket_mask;   // table buckets - 1 (a bucket = BK 8-byte keys, read with 16-byte loads)
    uint32_t bucket_shift;  // bucket = table_hash(key) >> bucket_shift
    uint32_t max_keys;      // largest build partition the table accepts
    uint64_t table_mul;     // odd multipliers of the table hash (derived from the table seed): low / high word
    uint32_t shift32;       // 32 - log2(buckets)
    uint32_t sentinel_part; // the one partition kEmptyKey hashes to: only there can a key equal the
                            // free-slot marker, every other partition runs without that check
    unsigned long long* matches;
    uint64_t* cta_times;  // 2 per CTA: build ns, probe ns
    uint32_t* oversize;   // optional: += the partitions skipped because their build side exceeds max_keys
};

// Hash of a key inside one partition's table. All keys of a partition share the partitioning hash's
// low bits, so the table uses an independent multiplicative (Fibonacci-style) hash: 5 instructions
// instead of a second 22-instruction XXH3. Like the reference's second XXHasher instance
// (src/main.cpp:215-217) it only influences speed, never the count.
__device__ __forceinline__ uint32_t table_bucket(uint64_t key, uint64_t mul, uint32_t shift) {
    return (uint32_t)(((key ^ (key >> 29)) * mul) >> shift);
}
// The same role with two 32-bit multiplies (multiply-shift over the key's halves): the probe loop of
// join_partitions is issue-bound (ncu: 72 % issue slots busy), a 64-bit IMAD chain costs 4 slots.
__device__ __forceinline__ uint32_t table_bucket32(uint64_t key, uint32_t mul_lo, uint32_t mul_hi, uint32_t shift32) {
    return ((uint32_t)key * mul_lo + (uint32_t)(key >> 32) * mul_hi) >> shift32;
}

template <int TPB, int BK>
__global__ void __launch_bounds__(TPB) join_partitions(JoinParams p) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint64_t* table = reinterpret_cast<uint64_t*>(smem_raw);
    __shared__ uint32_t has_empty_key;
    __shared__ unsigned long long block_count;
    const uint32_t nslots = (p.bucket_mask + 1) * BK;
    const int tid = threadIdx.x;

    // Slice of the partitioned probe relation owned by this CTA (the host did the division).
    const uint64_t lo = p.slice_len * blockIdx.x + min((uint64_t)blockIdx.x, p.slice_rem);
    const uint64_t hi = lo + p.slice_len + (blockIdx.x < p.slice_rem ? 1 : 0);
    if (tid == 0) block_count = 0;
    uint64_t build_ns = 0, probe_ns = 0;
    uint32_t count = 0;
    if (p.oversize)  // what count_oversize did in a launch of its own
        for (uint32_t i = blockIdx.x * TPB + tid; i < p.npart; i += gridDim.x * TPB)
            if (p.bounds_build[i + 1] - p.bounds_build[i] > p.max_keys) atomicAdd(p.oversize, 1u);

    if (lo < hi) {
        // last partition whose start is <= lo
        uint32_t a = 0, b = p.npart;
        while (b - a > 1) {
            const uint32_t m = (a + b) >> 1;
            if (p.bounds_probe[m] <= lo) a = m; else b = m;
        }
        for (uint32_t part = a; part < p.npart; ++part) {
            const uint64_t ps0 = p.bounds_probe[part];
            if (ps0 >= hi) break;
            const uint64_t s0 = max(lo, ps0), s1 = min(hi, p.bounds_probe[part + 1]);
            if (s0 >= s1) continue;
            const uint64_t r0 = p.bounds_build[part], r1 = p.bounds_build[part + 1];
            if (r1 == r0) continue;              // src/RadixCluster/HashJoin.hpp:273-276
            if (r1 - r0 > p.max_keys) continue;  // global-table path

            const uint64_t t0 = globaltimer_ns();
            cta_sync();  // previous partition's probes are done with the table
            {
                ulonglong2* t2 = reinterpret_cast<ulonglong2*>(table);
                for (uint32_t i = tid; i < nslots / 2; i += TPB) t2[i] = make_ulonglong2(kEmptyKey, kEmptyKey);
            }
            if (tid == 0) has_empty_key = 0;
            cta_sync();
            const bool careful = part == p.sentinel_part;  // only here can a key equal kEmptyKey
            const uint32_t mul_lo = (uint32_t)p.table_mul, mul_hi = (uint32_t)(p.table_mul >> 32);
            // ---- build: first free slot of the home bucket, overflowing into the next bucket.
            // Warp-converged: every lane runs the same number of loop trips (vote on `pending`),
            // so no per-lane loop ever splits the warp (see cta_sync()).
            for (uint64_t i0 = r0; i0 < r1; i0 += TPB) {
                const uint64_t i = i0 + tid;
                uint64_t key = 0;
                bool pending = i < r1;
                if (pending) {
                    key = __ldg(reinterpret_cast<const unsigned long long*>(p.build + i));
                    if (careful && key == kEmptyKey) {
                        has_empty_key = 1;
                        pending = false;
                    }
                }
                uint32_t slot = (table_bucket32(key, mul_lo, mul_hi, p.shift32) & p.bucket_mask) * BK;
                while (__any_sync(0xffffffffu, pending)) {
                    if (pending) {
                        unsigned long long cur = table[slot];
                        if (cur == kEmptyKey)
                            cur = atomicCAS(reinterpret_cast<unsigned long long*>(table + slot), kEmptyKey, key);
                        if (cur == kEmptyKey || cur == key) pending = false;  // duplicates collapse
                        else slot = (slot + 1) & (nslots - 1);
                    }
                }
            }
            cta_sync();
            const uint64_t t1 = globaltimer_ns();
            const uint32_t sentinel_hit = has_empty_key;

            // ---- probe: one bucket (BK keys) per step; a bucket whose last slot is free ends the
            // search (slots fill in order, LinearProbing.hpp:172-174 analogue). The first step is
            // peeled and branch-free: at the table's load factor nearly every search ends there, and
            // the converged loop only runs for the few that overflow. Full blocks of TPB * U probe
            // tuples skip the bounds checks; outside the sentinel partition no key can equal the
            // free-slot marker, so that test is skipped too (the loop is issue-bound: every
            // instruction removed here is time).
            constexpr int U = PHJ_JOIN_U;  // probe tuples (loads) in flight per thread
            auto probe_bucket = [&](uint64_t key, uint32_t bucket, bool& hit, bool& full) {
                if (BK == 2) {
                    const ulonglong2 k = *reinterpret_cast<const ulonglong2*>(table + bucket * 2);
                    hit = (k.x == key) | (k.y == key);
                    full = k.y != kEmptyKey;
                } else {
                    const ulonglong2* bk = reinterpret_cast<const ulonglong2*>(table + bucket * 4);
                    const ulonglong2 k01 = bk[0], k23 = bk[1];
                    hit = (k01.x == key) | (k01.y == key) | (k23.x == key) | (k23.y == key);
                    full = k23.y != kEmptyKey;
                }
            };
            auto overflow_walk = [&](uint64_t (&key)[U], uint32_t (&bucket)[U], bool (&pending)[U]) {
                auto any_pending = [&]() {
                    bool a = false;
#pragma unroll
                    for (int u = 0; u < U; ++u) a |= pending[u];
                    return a;
                };
                bool any = any_pending();
                while (__any_sync(0xffffffffu, any)) {
#pragma unroll
                    for (int u = 0; u < U; ++u)
                        if (pending[u]) {
                            bucket[u] = (bucket[u] + 1) & p.bucket_mask;
                            bool hit, full;
                            probe_bucket(key[u], bucket[u], hit, full);
                            count += hit;
                            pending[u] = !hit && full;
                        }
                    any = any_pending();
                }
            };
            uint64_t i0 = s0;
            if (!careful) {
                for (; i0 + (uint64_t)TPB * U <= s1; i0 += (uint64_t)TPB * U) {
                    uint64_t key[U];
                    uint32_t bucket[U];
                    bool pending[U];
#pragma unroll
                    for (int u = 0; u < U; ++u)
                        key[u] = ld_stream_u64(reinterpret_cast<const uint64_t*>(p.probe + i0 + (uint64_t)u * TPB + tid));
#pragma unroll
                    for (int u = 0; u < U; ++u) {
                        bucket[u] = table_bucket32(key[u], mul_lo, mul_hi, p.shift32) & p.bucket_mask;
                        bool hit, full;
                        probe_bucket(key[u], bucket[u], hit, full);
                        count += hit;
                        pending[u] = !hit && full;
                    }
                    overflow_walk(key, bucket, pending);
                }
            }
            for (; i0 < s1; i0 += (uint64_t)TPB * U) {  // tail block, and the sentinel partition
                uint64_t key[U];
                uint32_t bucket[U];
                bool pending[U];
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const uint64_t i = i0 + (uint64_t)u * TPB + tid;
                    pending[u] = i < s1;
                    key[u] = pending[u] ? ld_stream_u64(reinterpret_cast<const uint64_t*>(p.probe + i)) : 0;
                }
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    if (pending[u] && key[u] == kEmptyKey) {
                        count += sentinel_hit;
                        pending[u] = false;
                    }
                    bucket[u] = table_bucket32(key[u], mul_lo, mul_hi, p.shift32) & p.bucket_mask;
                    if (pending[u]) {
                        bool hit, full;
                        probe_bucket(key[u], bucket[u], hit, full);
                        count += hit;
                        pending[u] = !hit && full;
                    }
                }
                overflow_walk(key, bucket, pending);
            }
            const uint64_t t2 = globaltimer_ns();
            build_ns += t1 - t0;
            probe_ns += t2 - t1;
        }
    }
    cta_sync();
    // block reduction of the match count
    uint32_t c = count;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    if ((tid & 31) == 0 && c) atomicAdd(&block_count, (unsigned long long)c);
    cta_sync();
    if (tid == 0) {
        if (block_count) atomicAdd(p.matches, block_count);
        p.cta_times[2 * blockIdx.x] = build_ns;
        p.cta_times[2 * blockIdx.x + 1] = probe_ns;
    }
}

// =================================================================================================
// K5m  join_materialize: the joined table the reference declares but leaves empty
// (Table<JoinedTuple>{id, payloadA, payloadB}, src/Common/Table.hpp:27-33; Run() returns it empty,
// src/RadixCluster/HashJoin.hpp:226-227), filled with GetAll semantics
// (src/HashTables/LinearProbing.hpp:183-200): one row per (probe tuple, equal-key build tuple).
//
// Same slicing as join_partitions. Per partition the CTA copies up to cap_tuples build TUPLES into
// shared memory and links them into an open-addressing table of 32-bit slots holding tuple index + 1
// (0 = free), so duplicate build keys each keep their slot and no key value is reserved. A build
// side larger than cap_tuples is processed chunk by chunk against the same probe slice (block
// nested loop), so any partition size and any duplication is handled.
// WRITE = false only counts the rows, per CTA (sizes the output; a scan of the per-CTA counts gives
// every CTA its own output range). WRITE = true appends them: the lanes of a warp that found a
// match in the same probe step reserve their rows with ONE shared-memory atomicAdd on the CTA's
// cursor and write them side by side (24-byte rows, contiguous per warp step). (A single global
// cursor was tried first: 6 M same-address L2 atomics took 15 ms, measured.)
// =================================================================================================
struct MatParams {
    const ulonglong2* build;
    const ulonglong2* probe;
    const uint64_t* bounds_build;
    const uint64_t* bounds_probe;
    uint64_t n_probe;
    uint64_t slice_len, slice_rem;
    uint32_t npart;
    uint32_t cap_tuples;   // build tuples per table chunk
    uint32_t slot_mask;    // slots - 1, slots = power of two >= 2 * cap_tuples
    uint32_t slot_shift;   // slot = table_bucket(key) >> ... (64 - log2 slots)
    uint64_t table_mul;
    unsigned long long* cursor;  // WRITE = false: total row count
    unsigned long long* cta_rows;  // WRITE = false: out, rows of every CTA; WRITE = true: in, their
                                   // exclusive scan = where every CTA's rows start
    int64_t* out;                // rows of 3 x int64
    uint64_t out_cap;            // rows
};

template <int TPB, bool WRITE>
__global__ void __launch_bounds__(TPB) join_materialize(MatParams p) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    ulonglong2* tup = reinterpret_cast<ulonglong2*>(smem_raw);                  // [cap_tuples]
    uint32_t* slots = reinterpret_cast<uint32_t*>(tup + p.cap_tuples);           // [slot_mask + 1]
    __shared__ unsigned long long block_count;
    __shared__ uint64_t wstage_all[WRITE ? (TPB / 32) * 96 : 1];                 // 32 rows x 3 words per warp
    uint64_t* wstage = wstage_all + (WRITE ? (threadIdx.x >> 5) * 96 : 0);
    const int tid = threadIdx.x, lane = tid & 31;
    const uint32_t lt = lanemask_lt();

    const uint64_t lo = p.slice_len * blockIdx.x + min((uint64_t)blockIdx.x, p.slice_rem);
    const uint64_t hi = lo + p.slice_len + (blockIdx.x < p.slice_rem ? 1 : 0);
    if (tid == 0) block_count = WRITE ? p.cta_rows[blockIdx.x] : 0;  // WRITE: this CTA's append cursor
    unsigned long long count = 0;

    if (lo < hi) {
        uint32_t a = 0, b = p.npart;
        while (b - a > 1) {
            const uint32_t m = (a + b) >> 1;
            if (p.bounds_probe[m] <= lo) a = m; else b = m;
        }
        for (uint32_t part = a; part < p.npart; ++part) {
            const uint64_t ps0 = p.bounds_probe[part];
            if (ps0 >= hi) break;
            const uint64_t s0 = max(lo, ps0), s1 = min(hi, p.bounds_probe[part + 1]);
            if (s0 >= s1) continue;
            const uint64_t r0 = p.bounds_build[part], r1 = p.bounds_build[part + 1];
            for (uint64_t c0 = r0; c0 < r1; c0 += p.cap_tuples) {
                const uint32_t nb = (uint32_t)min((uint64_t)p.cap_tuples, r1 - c0);
                cta_sync();  // the previous chunk's probes are done with the table
                for (uint32_t i = tid; i <= p.slot_mask; i += TPB) slots[i] = 0;
                for (uint32_t i = tid; i < nb; i += TPB) tup[i] = p.build[c0 + i];
                cta_sync();
                // ---- build: every tuple claims the first free slot of its probe sequence ----
                for (uint32_t i0 = 0; i0 < nb; i0 += TPB) {
                    const uint32_t i = i0 + tid;
                    bool pending = i < nb;
                    uint32_t slot = pending ? table_bucket(tup[i].x, p.table_mul, p.slot_shift) & p.slot_mask : 0;
                    while (__any_sync(0xffffffffu, pending)) {
                        if (pending) {
                            uint32_t cur = slots[slot];
                            if (cur == 0) cur = atomicCAS(&slots[slot], 0u, i + 1);
                            if (cur == 0) pending = false;
                            else slot = (slot + 1) & p.slot_mask;
                        }
                    }
                }
                cta_sync();
                // ---- probe: walk the sequence to the first free slot, every equal key is a row.
                // U probe tuples per thread are in flight (independent shared-memory chains).
                constexpr int U = 4;
                for (uint64_t i0 = s0; i0 < s1; i0 += (uint64_t)TPB * U) {
                    ulonglong2 s[U];
                    uint32_t slot[U];
                    bool pending[U];
#pragma unroll
                    for (int u = 0; u < U; ++u) {
                        const uint64_t i = i0 + (uint64_t)u * TPB + tid;
                        pending[u] = i < s1;
                        s[u] = make_ulonglong2(0, 0);
                        if (pending[u]) {
                            if (WRITE) s[u] = ld_stream_v2(p.probe + i);
                            else s[u].x = ld_stream_u64(reinterpret_cast<const uint64_t*>(p.probe + i));
                        }
                    }
#pragma unroll
                    for (int u = 0; u < U; ++u) slot[u] = table_bucket(s[u].x, p.table_mul, p.slot_shift) & p.slot_mask;
                    while (__any_sync(0xffffffffu, pending[0] | pending[1] | pending[2] | pending[3])) {
#pragma unroll
                        for (int u = 0; u < U; ++u) {
                            bool hit = false;
                            ulonglong2 r = make_ulonglong2(0, 0);
                            if (pending[u]) {
                                const uint32_t cur = slots[slot[u]];
                                if (cur == 0) {
                                    pending[u] = false;
                                } else {
                                    if (WRITE) r = tup[cur - 1];
                                    else r.x = tup[cur - 1].x;
                                    hit = r.x == s[u].x;
                                    slot[u] = (slot[u] + 1) & p.slot_mask;
                                }
                            }
                            if (WRITE) {
                                const uint32_t hits = __ballot_sync(0xffffffffu, hit);
                                if (hits) {
                                    unsigned long long base = 0;
                                    if (lane == __ffs(hits) - 1)
                                        base = atomicAdd(&block_count, (unsigned long long)__popc(hits));
                                    base = __shfl_sync(0xffffffffu, base, __ffs(hits) - 1);
                                    // rows -> the warp's staging words, then consecutive lanes store
                                    // consecutive 8-byte words: full 32-byte sectors instead of three
                                    // strided partial-sector stores per row
                                    const uint32_t nh = __popc(hits);
                                    if (hit) {
                                        uint64_t* w = wstage + 3 * __popc(hits & lt);
                                        w[0] = s[u].x;
                                        w[1] = r.y;
                                        w[2] = s[u].y;
                                    }
                                    __syncwarp();
                                    if (base + nh <= p.out_cap) {
                                        uint64_t* o = reinterpret_cast<uint64_t*>(p.out) + base * 3;
#pragma unroll
                                        for (int k = 0; k < 3; ++k)
                                            if (lane + 32 * k < 3 * nh) o[lane + 32 * k] = wstage[lane + 32 * k];
                                    }
                                    __syncwarp();
                                }
                            } else {
                                count += hit;
                            }
                        }
                    }
                }
            }
        }
    }
    if (!WRITE) {
        cta_sync();
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) count += __shfl_xor_sync(0xffffffffu, count, o);
        if (lane == 0 && count) atomicAdd(&block_count, count);
        cta_sync();
        if (tid == 0) {
            p.cta_rows[blockIdx.x] = block_count;
            if (block_count) atomicAdd(p.cursor, block_count);
        }
    }
}

// Exclusive scan of the per-CTA row counts (<= 8192 CTAs), in place. One CTA.
__global__ void __launch_bounds__(1024) scan_cta_rows(unsigned long long* rows, uint32_t n) {
    __shared__ uint64_t sh[33];
    uint64_t v[8], s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const uint32_t idx = threadIdx.x * 8 + i;
        v[i] = idx < n ? rows[idx] : 0;
        s += v[i];
    }
    uint64_t excl = block_excl_scan_u64(s, sh, nullptr);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const uint32_t idx = threadIdx.x * 8 + i;
        if (idx < n) rows[idx] = excl;
        excl += v[i];
    }
}

// Counts the partitions join_partitions skipped (build side larger than the smem table).
__global__ void count_oversize(const uint64_t* __restrict__ bounds_build, uint32_t npart,
                               uint32_t max_keys, uint32_t* __restrict__ oversize) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < npart && bounds_build[i + 1] - bounds_build[i] > max_keys) atomicAdd(oversize, 1u);
}

// =================================================================================================
// K4g/K5g  pt_build / pt_probe: per-partition tables in GLOBAL memory sized for the L2
// (PHJ_FLAG_L2_TABLES, and the local join of the multi-GPU path). With few, large partitions (one
// radix pass of 6 bits: 156 K build keys each at 10 M x 200 M) a partition's table (32-byte buckets of
// four keys, load <= 0.5) is a few MB; the probe relation is walked in partition order by all CTAs
// together, so only the tables of the partitions being probed are hot and they stay in the 126 MB L2.
// Unlike gt_build / gt_probe (the fallback for oversize partitions) the partition comes from the
// tuple's POSITION (the boundaries), so the table hash is the 5-instruction multiply-shift of
// join_partitions instead of a second XXH3, and there is no per-tuple selection test. Measured
// (round 2, 10 M x 200 M uniform): one 6-bit pass + this join 3.41 ms against 3.95 ms for two passes +
// shared-memory tables -- the second partitioning pass (1.36 ms + planning) costs more than probing
// out of L2 (0.37 + 1.14 ms) instead of shared memory (0.62 ms).
// Everything the kernels need to know about sizes comes from the device-resident boundaries
// (bounds[npart] is the relation's size), so the multi-GPU path can enqueue them before the host
// knows how many tuples the shuffle delivered. Boundaries may be absolute positions inside a larger
// window: the probe side is sliced over [bounds_probe[0], bounds_probe[npart]).
// Partitions whose build side exceeds max_keys are skipped and counted, as in join_partitions.
// =================================================================================================
struct PtParams {
    const ulonglong2* build;       // partitioned R (or the build window of the shuffle)
    const ulonglong2* probe;       // partitioned S (or one chunk's region of the probe window)
    const uint64_t* bounds_build;  // npart + 1, device
    const uint64_t* bounds_probe;  // npart + 1, device
    uint32_t npart;
    uint32_t max_keys;             // largest build partition a region accepts
    uint64_t table_mul;            // odd multipliers of the table hash: low / high word
    uint32_t sentinel_part;        // the one partition kEmptyKey hashes to (0xffffffff: none here)
    uint32_t region_shift32;       // 32 - log2(region_buckets)
    uint32_t region_buckets;       // a power of two
    uint64_t* table;               // npart regions of region_buckets buckets of 4 keys
    uint32_t* flags;               // [0]: the sentinel partition's build side contains kEmptyKey
    unsigned long long* matches;
};

// First partition whose range [bounds[a], bounds[a + 1]) can contain position `pos`.
__device__ __forceinline__ uint32_t partition_of(const uint64_t* __restrict__ bounds, uint32_t npart, uint64_t pos) {
    uint32_t a = 0, b = npart;
    while (b - a > 1) {
        const uint32_t m = (a + b) >> 1;
        if (bounds[m] <= pos) a = m;
        else b = m;
    }
    return a;
}

// Both kernels walk their relation in TILES dealt round-robin to the CTAs (tile t goes to CTA t mod grid),
// so that all resident CTAs work on neighbouring tiles: at any moment only the one or two partitions
// under that moving front have their tables touched, and those stay in L2 whatever the total size of
// all tables is. (First version: one contiguous slice per CTA -- every partition's table was live at
// once, 268 MB of them at 10 M build keys, and two thirds of the lookups went to DRAM at 64 bytes each.)
template <int TPB>
__global__ void __launch_bounds__(TPB) pt_build(PtParams p) {
    constexpr int U = 2;
    constexpr uint64_t TILE = (uint64_t)TPB * U;
    const uint64_t first = p.bounds_build[0], end = p.bounds_build[p.npart];
    const uint32_t mul_lo = (uint32_t)p.table_mul, mul_hi = (uint32_t)(p.table_mul >> 32);
    const uint32_t bmask = p.region_buckets - 1;
    for (uint64_t t0 = first + (uint64_t)blockIdx.x * TILE; t0 < end; t0 += (uint64_t)gridDim.x * TILE) {
        const uint32_t part0 = partition_of(p.bounds_build, p.npart, t0);
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const uint64_t i = t0 + (uint64_t)u * TPB + threadIdx.x;
            bool pending = i < end;
            uint32_t part = part0;
            while (pending && i >= p.bounds_build[part + 1]) ++part;  // a tile may straddle (empty) partitions
            uint64_t key = 0;
            if (pending) {
                if (p.bounds_build[part + 1] - p.bounds_build[part] > p.max_keys) {
                    pending = false;  // oversize: counted by the caller, joined elsewhere
                } else {
                    key = ld_stream_u64(reinterpret_cast<const uint64_t*>(p.build + i));
                    if (part == p.sentinel_part && key == kEmptyKey) {
                        p.flags[0] = 1;
                        pending = false;
                    }
                }
            }
            uint64_t* __restrict__ region = p.table + (uint64_t)part * p.region_buckets * 4;
            uint32_t bucket = table_bucket32(key, mul_lo, mul_hi, p.region_shift32) & bmask, s = 0;
            // warp-converged insert loop (see cta_sync()): first free slot of the home bucket, then the next bucket
            while (__any_sync(0xffffffffu, pending)) {
                if (pending) {
                    unsigned long long* slot = reinterpret_cast<unsigned long long*>(region + (uint64_t)bucket * 4 + s);
                    unsigned long long cur = *slot;
                    if (cur == kEmptyKey) cur = atomicCAS(slot, kEmptyKey, key);
                    if (cur == kEmptyKey || cur == key) {
                        pending = false;  // duplicates collapse
                    } else if (++s == 4) {
                        s = 0;
                        bucket = (bucket + 1) & bmask;
                    }
                }
            }
        }
    }
}

// (A/B, ncu r02d: one contiguous slice per CTA instead -- all 64 tables live at once -- had an L2 hit
// rate of 12 % and read 19.4 GB from DRAM, 3.34 ms; the round-robin tiles read 3.6 GB, 1.44 ms.)
// (Tried in round 2, ncu r02l: the probe tuples streamed into shared memory by TMA bulk loads -- double buffer,
// one mbarrier per buffer, cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes -- so that only the
// bucket reads go through the L1: l1tex sectors 518 M -> 216 M as intended, but 3.66 ms instead of 1.13 ms. The
// block-wide barrier that hands a buffer back to the loader makes every tile wait for the slowest of its 1024
// bucket reads (long_scoreboard 70 %, issue slots 21 % busy), and half of those miss the L2 at uniform keys.
// It needs per-warp "empty" barriers and a deeper ring, at which point shared memory caps the CTAs per SM
// below what hides the lookup latency. The register-staged stream below stays.)
// (A/B, ncu r02d: one contiguous slice per CTA instead -- all 64 tables live at once -- had an L2 hit
// rate of 12 % and read 19.4 GB from DRAM, 3.34 ms; the round-robin tiles read 3.6 GB, 1.44 ms.)
template <int TPB>
__global__ void __launch_bounds__(TPB) pt_probe(PtParams p) {
    __shared__ unsigned long long block_count;
    if (threadIdx.x == 0) block_count = 0;
    constexpr int U = 4;
    constexpr uint64_t TILE = (uint64_t)TPB * U;
    const uint64_t first = p.bounds_probe[0], end = p.bounds_probe[p.npart];
    const uint32_t mul_lo = (uint32_t)p.table_mul, mul_hi = (uint32_t)(p.table_mul >> 32);
    const uint32_t bmask = p.region_buckets - 1;
    const uint32_t sentinel_hit = p.flags[0];
    uint32_t count = 0;
    const uint64_t lo = first + (uint64_t)blockIdx.x * TILE;
    uint32_t part0 = lo < end ? partition_of(p.bounds_probe, p.npart, lo) : 0;
    for (uint64_t t0 = lo; t0 < end; t0 += (uint64_t)gridDim.x * TILE) {
        uint64_t key[U];
        bool pending[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const uint64_t i = t0 + (uint64_t)u * TPB + threadIdx.x;
            pending[u] = i < end;
            key[u] = pending[u] ? ld_stream_u64(reinterpret_cast<const uint64_t*>(p.probe + i)) : 0;
        }
        // the tile's partition(s): the front only moves forward; nearly always the tile lies in one
        // partition, a tile that straddles a boundary resolves per tuple
        while (t0 >= p.bounds_probe[part0 + 1]) ++part0;
        const bool uniform = min(t0 + TILE, end) <= p.bounds_probe[part0 + 1];
        uint32_t part[U], bucket[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            part[u] = part0;
            if (!uniform) {
                const uint64_t i = t0 + (uint64_t)u * TPB + threadIdx.x;
                while (pending[u] && i >= p.bounds_probe[part[u] + 1]) ++part[u];
            }
            const uint64_t nb = p.bounds_build[part[u] + 1] - p.bounds_build[part[u]];
            if (nb == 0 || nb > p.max_keys) pending[u] = false;  // src/RadixCluster/HashJoin.hpp:273-276 / oversize
            if (part[u] == p.sentinel_part && pending[u] && key[u] == kEmptyKey) {
                count += sentinel_hit;
                pending[u] = false;
            }
            bucket[u] = table_bucket32(key[u], mul_lo, mul_hi, p.region_shift32) & bmask;
        }
        // one 32-byte bucket per step; a bucket whose last slot is free ends the search
        while (__any_sync(0xffffffffu, pending[0] | pending[1] | pending[2] | pending[3])) {
#pragma unroll
            for (int u = 0; u < U; ++u) {
                if (pending[u]) {
                    const Bucket4 b = ld_bucket4(p.table + ((uint64_t)part[u] * p.region_buckets + bucket[u]) * 4);
                    if (b.k0 == key[u] || b.k1 == key[u] || b.k2 == key[u] || b.k3 == key[u]) {
                        ++count;
                        pending[u] = false;
                    } else if (b.k3 == kEmptyKey) {
                        pending[u] = false;
                    } else {
                        bucket[u] = (bucket[u] + 1) & bmask;
                    }
                }
            }
        }
    }
    cta_sync();
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) count += __shfl_xor_sync(0xffffffffu, count, o);
    if ((threadIdx.x & 31) == 0 && count) atomicAdd(&block_count, (unsigned long long)count);
    cta_sync();
    if (threadIdx.x == 0 && block_count) atomicAdd(p.matches, block_count);
}

// =================================================================================================
// K6/K7  global open-addressing table: the no-partitioning join, and the fallback for oversize
// radix partitions. Buckets are 32-byte sectors of four 8-byte keys; a key hashes to a bucket,
// takes the first free slot (atomicCAS) and overflows into the next bucket -- the GPU analogue of
// the reference's 64-byte, 3-slot LinearProbing buckets (src/HashTables/LinearProbing.hpp:22-83,
// 114-134). A probe reads one whole bucket with two 16-byte loads, so the common case costs one
// DRAM sector; it stops at the first bucket that has a free slot (LinearProbing.hpp:172-174).
// =================================================================================================
struct GtParams {
    const ulonglong2* rel;     // R for build, S for probe
    uint64_t n;
    uint64_t* table;           // nbuckets * 4 keys
    uint64_t bucket_mask;      // nbuckets - 1
    uint32_t hash_shift;
    uint32_t select;           // 0: every tuple; 1: only tuples of oversize partitions
    const uint64_t* bounds_build;  // select == 1: partition sizes of R
    uint32_t max_keys;
    DigitFn part_fn;           // select == 1: tuple -> partition id (shift 0, mask = all)
    // pre-partitioned relations: partition = (pass-1 digit - pre_first) * 2^pre_b2 + low bits, with
    // the pass-1 digit = (hash >> pre_b2) & pre_space_mask (pre_space_mask == 0: not pre-partitioned)
    uint64_t pre_space_mask;
    uint32_t pre_first, pre_b2;
    uint32_t region_bits;      // select == 1: the table is cut into 2^region_bits regions, one per
                               // group of partitions, so a partition's keys share a small, L2-sized
                               // piece of the table and the partition-ordered probe stays in L2
    uint32_t region_shift;     // region = partition >> region_shift
    uint64_t local_mask;       // buckets per region - 1
    HashParams hp;
    uint32_t* flags;           // [0]: build side contains kEmptyKey
    unsigned long long* matches;
};

__global__ void gt_clear(uint64_t* __restrict__ table, uint64_t nkeys) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    ulonglong2* t2 = reinterpret_cast<ulonglong2*>(table);
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nkeys / 2; i += stride)
        t2[i] = make_ulonglong2(kEmptyKey, kEmptyKey);
}

template <bool POW2>
__device__ __forceinline__ uint64_t gt_part(const GtParams& p, uint64_t h) {
    if (p.pre_space_mask)
        return ((((h >> p.pre_b2) & p.pre_space_mask) - p.pre_first) << p.pre_b2) | (h & ((1ull << p.pre_b2) - 1));
    return POW2 ? (h & p.part_fn.pmask) : (h % p.part_fn.modulus);
}

template <bool POW2>
__device__ __forceinline__ uint64_t gt_bucket(const GtParams& p, uint64_t h) {
    const uint64_t local = h >> p.hash_shift;
    if (!p.select || p.region_bits == 0) return local & p.bucket_mask;
    const uint64_t part = gt_part<POW2>(p, h);
    return (((part >> p.region_shift) & ((1ull << p.region_bits) - 1)) * (p.local_mask + 1)) | (local & p.local_mask);
}

template <bool POW2>
__device__ __forceinline__ bool gt_selected(const GtParams& p, uint64_t h) {
    if (!p.select) return true;
    const uint64_t part = gt_part<POW2>(p, h);  // every tuple of a bound relation lies in a bound parent
    return p.bounds_build[part + 1] - p.bounds_build[part] > p.max_keys;
}

template <int HASH, bool POW2>
__global__ void __launch_bounds__(256) gt_build(GtParams p) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    // uniform trip count per warp; per-lane work is predicated (warp-converged loops, see cta_sync())
    for (uint64_t base = (uint64_t)blockIdx.x * blockDim.x; base < p.n; base += stride) {
        const uint64_t i = base + threadIdx.x;
        bool pending = i < p.n;
        const uint64_t key = pending ? ld_stream_u64(reinterpret_cast<const uint64_t*>(p.rel + i)) : 0;
        const uint64_t h = hash_key<HASH>(key, p.hp);
        if (pending && !gt_selected<POW2>(p, h)) pending = false;
        if (pending && key == kEmptyKey) {
            p.flags[0] = 1;
            pending = false;
        }
        uint64_t bucket = gt_bucket<POW2>(p, h);
        uint32_t s = 0;  // slot inside the bucket
        while (__any_sync(0xffffffffu, pending)) {
            if (pending) {
                unsigned long long* slot = reinterpret_cast<unsigned long long*>(p.table + bucket * 4 + s);
                unsigned long long cur = *slot;
                if (cur == kEmptyKey) cur = atomicCAS(slot, kEmptyKey, key);
                if (cur == kEmptyKey || cur == key) {
                    pending = false;
                } else if (++s == 4) {
                    s = 0;
                    bucket = (bucket + 1) & p.bucket_mask;
                }
            }
        }
    }
}

template <int HASH, bool POW2>
__global__ void __launch_bounds__(256) gt_probe(GtParams p) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    const uint32_t sentinel_hit = p.flags[0];
    uint32_t count = 0;
    constexpr int U = 4;
    for (uint64_t base = (uint64_t)blockIdx.x * blockDim.x; base < p.n; base += stride * U) {
        uint64_t key[U], bucket[U];
        bool pending[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const uint64_t i = base + (uint64_t)u * stride + threadIdx.x;
            pending[u] = i < p.n;
            key[u] = pending[u] ? ld_stream_u64(reinterpret_cast<const uint64_t*>(p.rel + i)) : 0;
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const uint64_t h = hash_key<HASH>(key[u], p.hp);
            if (pending[u] && !gt_selected<POW2>(p, h)) pending[u] = false;
            if (pending[u] && key[u] == kEmptyKey) {
                count += sentinel_hit;
                pending[u] = false;
            }
            bucket[u] = gt_bucket<POW2>(p, h);
        }
        while (__any_sync(0xffffffffu, pending[0] | pending[1] | pending[2] | pending[3])) {
#pragma unroll
            for (int u = 0; u < U; ++u) {
                if (pending[u]) {
                    const Bucket4 b = ld_bucket4(p.table + bucket[u] * 4);
                    if (b.k0 == key[u] || b.k1 == key[u] || b.k2 == key[u] || b.k3 == key[u]) {
                        ++count;
                        pending[u] = false;
                    } else if (b.k3 == kEmptyKey) {
                        pending[u] = false;  // slots fill in order: last one free => bucket not full
                    } else {
                        bucket[u] = (bucket[u] + 1) & p.bucket_mask;
                    }
                }
            }
        }
    }
    __shared__ unsigned long long block_count;
    if (threadIdx.x == 0) block_count = 0;
    cta_sync();
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) count += __shfl_xor_sync(0xffffffffu, count, o);
    if ((threadIdx.x & 31) == 0 && count) atomicAdd(&block_count, (unsigned long long)count);
    cta_sync();
    if (threadIdx.x == 0 && block_count) atomicAdd(p.matches, block_count);
}

// A/B for the north star's "probed by cooperative warps using ballot and match primitives": four lanes
// share one probe -- the owner lane's key and bucket are broadcast with shuffles, every lane of the group
// reads ONE 8-byte slot of the 32-byte bucket (the group's four loads coalesce into one sector), hit and
// "last slot free" are found with a ballot over the group. Same sectors as the per-thread probe, four
// times the load instructions plus shuffles and a vote per step. PHJ_FLAG_COOP_PROBE selects it.
template <int HASH, bool POW2>
__global__ void __launch_bounds__(256) gt_probe_coop(GtParams p) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    const uint32_t sentinel_hit = p.flags[0];
    const int lane = threadIdx.x & 31, sub = lane & 3, gbase = lane & ~3;
    uint32_t count = 0;
    for (uint64_t base = (uint64_t)blockIdx.x * blockDim.x; base < p.n; base += stride) {
        const uint64_t i = base + threadIdx.x;
        bool mine = i < p.n;
        const uint64_t key = mine ? ld_stream_u64(reinterpret_cast<const uint64_t*>(p.rel + i)) : 0;
        const uint64_t h = hash_key<HASH>(key, p.hp);
        if (mine && !gt_selected<POW2>(p, h)) mine = false;
        if (mine && key == kEmptyKey) {
            count += sentinel_hit;
            mine = false;
        }
        const uint64_t my_bucket = gt_bucket<POW2>(p, h);
#pragma unroll
        for (int r = 0; r < 4; ++r) {  // the group's four probes, one after the other
            const uint64_t k = __shfl_sync(0xffffffffu, key, gbase + r);
            uint64_t bucket = __shfl_sync(0xffffffffu, my_bucket, gbase + r);
            bool pending = __shfl_sync(0xffffffffu, (int)mine, gbase + r) != 0;
            while (__any_sync(0xffffffffu, pending)) {
                uint64_t slot = kEmptyKey;
                if (pending) slot = __ldg(reinterpret_cast<const unsigned long long*>(p.table + bucket * 4 + sub));
                const uint32_t hits = __ballot_sync(0xffffffffu, pending && slot == k);
                const uint32_t frees = __ballot_sync(0xffffffffu, pending && sub == 3 && slot == kEmptyKey);
                if (pending) {
                    const uint32_t gmask = 0xfu << gbase;
                    if (hits & gmask) {
                        if (sub == r) ++count;  // counted once, by the lane that owns the key
                        pending = false;
                    } else if (frees & gmask) {
                        pending = false;
                    } else {
                        bucket = (bucket + 1) & p.bucket_mask;
                    }
                }
            }
        }
    }
    __shared__ unsigned long long block_count;
    if (threadIdx.x == 0) block_count = 0;
    cta_sync();
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) count += __shfl_xor_sync(0xffffffffu, count, o);
    if ((threadIdx.x & 31) == 0 && count) atomicAdd(&block_count, (unsigned long long)count);
    cta_sync();
    if (threadIdx.x == 0 && block_count) atomicAdd(p.matches, block_count);
}

// =================================================================================================
// K6c/K7c  bucket-chained global table: the GPU counterpart of SeparateChainingHashTable
// (src/HashTables/SeparateChaining.hpp:143-277). A key hashes to one of `nheads` chain heads; a
// bucket is one 32-byte sector {3 keys, next, count} -- the reference's 3-slot bucket
// (SeparateChaining.hpp:22-101) shrunk from 64 to 32 bytes because only keys are stored; the first
// bucket of every chain is preallocated (bucket i belongs to head i, :176-181), overflow buckets
// come from a bump allocator (BucketAllocator, :103-135) and are pushed at the chain HEAD (:226-237).
// Insert: claim a slot of the head bucket with atomicAdd on its count; when it is full, ONE inserter
// wins the head's lock (atomicCAS to kChainLocked), links a fresh bucket holding its key and
// publishes it; the others retry on the new head. All loops are warp-converged (see cta_sync()).
// Duplicates keep their own slots (GetAll semantics); the count-only probe stops at the first hit.
// =================================================================================================
constexpr uint32_t kChainLocked = 0xffffffffu;

struct __align__(32) ChainBucket {
    uint64_t key[3];
    uint32_t next;   // bucket index + 1 of the next (older) bucket, 0 = end of chain
    uint32_t count;  // slots claimed (may exceed 3 transiently; valid slots = min(count, 3))
};

struct CtParams {
    const ulonglong2* rel;
    uint64_t n;
    uint32_t* heads;        // [nheads]: bucket index + 1 of the newest bucket of the chain
    ChainBucket* buckets;   // [pool]
    uint32_t nheads_mask;
    uint32_t hash_shift;
    uint32_t pool;          // buckets available
    uint32_t* cursor;       // bump allocator: next free bucket (starts at nheads)
    uint32_t* overflow;     // set when the allocator ran out ("BucketAllocator exceeded its limit.")
    HashParams hp;
    unsigned long long* matches;
};

__global__ void ct_init(CtParams p) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i <= p.nheads_mask; i += stride) {
        p.heads[i] = (uint32_t)i + 1;
        p.buckets[i].next = 0;
        p.buckets[i].count = 0;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        *p.cursor = p.nheads_mask + 1;
        *p.overflow = 0;
    }
}

template <int HASH>
__global__ void __launch_bounds__(256) ct_build(CtParams p) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t base = (uint64_t)blockIdx.x * blockDim.x; base < p.n; base += stride) {
        const uint64_t i = base + threadIdx.x;
        bool pending = i < p.n;
        const uint64_t key = pending ? ld_stream_u64(reinterpret_cast<const uint64_t*>(p.rel + i)) : 0;
        const uint32_t h = (uint32_t)(hash_key<HASH>(key, p.hp) >> p.hash_shift) & p.nheads_mask;
        volatile uint32_t* head = p.heads + h;
        while (__any_sync(0xffffffffu, pending)) {
            if (pending) {
                const uint32_t hb = *head;
                if (hb != kChainLocked) {
                    ChainBucket* b = p.buckets + (hb - 1);
                    uint32_t c = *reinterpret_cast<volatile uint32_t*>(&b->count);
                    if (c < 3) c = atomicAdd(&b->count, 1u);
                    if (c < 3) {
                        b->key[c] = key;
                        pending = false;
                    } else if (atomicCAS(p.heads + h, hb, kChainLocked) == hb) {
                        // this inserter extends the chain: new bucket at the head, holding its key
                        const uint32_t nb = atomicAdd(p.cursor, 1u);
                        if (nb < p.pool) {
                            ChainBucket* fresh = p.buckets + nb;
                            fresh->key[0] = key;
                            fresh->next = hb;
                            fresh->count = 1;
                            __threadfence();
                            *head = nb + 1;
                        } else {
                            *p.overflow = 1;  // allocator exhausted: the key is dropped, the join fails
                            __threadfence();
                            *head = hb;
                        }
                        pending = false;
                    }
                }
            }
        }
    }
}

template <int HASH>
__global__ void __launch_bounds__(256) ct_probe(CtParams p) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    uint32_t count = 0;
    constexpr int U = 4;
    for (uint64_t base = (uint64_t)blockIdx.x * blockDim.x; base < p.n; base += stride * U) {
        uint64_t key[U];
        uint32_t b[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const uint64_t i = base + (uint64_t)u * stride + threadIdx.x;
            const bool valid = i < p.n;
            key[u] = valid ? ld_stream_u64(reinterpret_cast<const uint64_t*>(p.rel + i)) : 0;
            b[u] = 0;
            if (valid) b[u] = __ldg(p.heads + ((uint32_t)(hash_key<HASH>(key[u], p.hp) >> p.hash_shift) & p.nheads_mask));
        }
        while (__any_sync(0xffffffffu, (b[0] | b[1] | b[2] | b[3]) != 0)) {
#pragma unroll
            for (int u = 0; u < U; ++u) {
                if (b[u]) {
                    const ulonglong2* q = reinterpret_cast<const ulonglong2*>(p.buckets + (b[u] - 1));
                    const ulonglong2 k01 = __ldg(q), k2m = __ldg(q + 1);
                    const uint32_t next = (uint32_t)k2m.y, cnt = (uint32_t)(k2m.y >> 32);
                    const bool hit = (cnt > 0 && k01.x == key[u]) | (cnt > 1 && k01.y == key[u]) |
                                     (cnt > 2 && k2m.x == key[u]);
                    count += hit;
                    b[u] = hit ? 0 : next;
                }
            }
        }
    }
    __shared__ unsigned long long block_count;
    if (threadIdx.x == 0) block_count = 0;
    cta_sync();
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) count += __shfl_xor_sync(0xffffffffu, count, o);
    if ((threadIdx.x & 31) == 0 && count) atomicAdd(&block_count, (unsigned long long)count);
    cta_sync();
    if (threadIdx.x == 0 && block_count) atomicAdd(p.matches, block_count);
}

// Multi-GPU split in K row chunks: where chunk c's tuples of digit d start in the (virtual) split
// output, starts[(rel * ndigits + d) * (K + 1) + c]; entry K is the end of the digit. Chunk c of a
// relation is the segment range [first_seg[rel][c], first_seg[rel][c + 1]).
// The sharded join counts its row shard piece by piece (piece 0 = the build relation, piece 1 + c = chunk c of the
// probe relation; with enough GPUs the count of chunk c + 1 runs while chunk c travels). One CTA per (split digit,
// piece) scans the digit's counters over THAT piece's segments only, so the scatter cursors of a piece are 0-based
// per digit, and writes the digit's total to sizes[(rel * ndig + d) * (K + 1) + c] (the build relation: c = 0).
struct ChunkScanParams {
    const uint32_t* counts;
    uint64_t* cursors;
    uint64_t* sizes;
    uint32_t cnt_base[2], nseg_rel[2];  // a relation's counter block: [digit][segment]
    uint32_t ndig, K;
    uint32_t piece_first;               // blockIdx.y = 0 is this piece
    uint32_t seg_first[kMaxPieces], seg_count[kMaxPieces];  // per piece: its segments inside its relation
};
__global__ void __launch_bounds__(256) chunk_scan(ChunkScanParams p) {
    __shared__ uint64_t sh[33];
    const uint32_t d = blockIdx.x, piece = p.piece_first + blockIdx.y;
    const uint32_t rel = piece ? 1 : 0, c = piece ? piece - 1 : 0, n = p.seg_count[piece];
    const uint64_t idx0 = p.cnt_base[rel] + (uint64_t)d * p.nseg_rel[rel] + p.seg_first[piece];
    uint64_t carry = 0;
    for (uint32_t i0 = 0; i0 < n; i0 += 256) {
        const uint32_t i = i0 + threadIdx.x;
        const uint64_t v = i < n ? p.counts[idx0 + i] : 0;
        uint64_t total;
        const uint64_t excl = block_excl_scan_u64(v, sh, &total);
        if (i < n) p.cursors[idx0 + i] = carry + excl;
        carry += total;
    }
    if (threadIdx.x == 0) p.sizes[((uint64_t)rel * p.ndig + d) * (p.K + 1) + c] = carry;
}

struct SplitStartsParams {
    const uint64_t* cursors;
    uint64_t* starts;
    uint32_t cnt_base[2], nseg[2];
    uint64_t bias[2], n[2];
    uint32_t ndigits, nchunks;
    uint32_t first_seg[2][17];  // up to 16 chunks
};
__global__ void split_starts(SplitStartsParams p) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t per_rel = p.ndigits * (p.nchunks + 1);
    if (i >= 2 * per_rel) return;
    const int rel = i / per_rel;
    const uint32_t d = (i % per_rel) / (p.nchunks + 1), c = i % (p.nchunks + 1);
    uint64_t v = 0;
    if (p.nseg[rel]) {
        const uint32_t s = c < p.nchunks ? p.first_seg[rel][c] : p.nseg[rel];
        if (s < p.nseg[rel]) v = p.cursors[p.cnt_base[rel] + (uint64_t)d * p.nseg[rel] + s] - p.bias[rel];
        else if (d + 1 < p.ndigits) v = p.cursors[p.cnt_base[rel] + (uint64_t)(d + 1) * p.nseg[rel]] - p.bias[rel];
        else v = p.n[rel];
    }
    p.starts[i] = v;
}

// =================================================================================================
// Device-side input generators: the reference's Sequential and Zipf generators
// (src/DataGenerator/Sequential.cpp:20-25, src/DataGenerator/Zipf.cpp:14-56,80-92) with the same
// batch / seed scheme as the host generator (batch b = an LCG seeded base_seed + b,
// src/Common/Random.cpp:9-30), one thread per batch. For scale runs: the arithmetic is the
// reference's, but pow() here is CUDA's, not glibc's, so a sample can differ from the host
// generator's where an intermediate lands within an ulp of an integer -- parity runs keep the host
// generator (phj_fill_zipf) and upload.
// =================================================================================================
__global__ void gen_sequential(ulonglong2* __restrict__ out, uint64_t n, int64_t start) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        out[i] = make_ulonglong2((uint64_t)(start + (int64_t)i), i);
}

struct ZipfGenParams {
    ulonglong2* out;
    uint64_t n, batches, batch;  // batch = n / batches tuples, the last batch takes the remainder
    double alpha, sd, norm;      // as set up by Zipf::generate (Zipf.cpp:22-28)
    int64_t base_seed, correction;
};

__device__ __forceinline__ double lcg_next(int64_t& state) {
    const int64_t a = 16807, m = 2147483647, q = 127773, r = 2836;
    const int64_t t = a * (state % q) - r * (state / q);
    state = t > 0 ? t : t + m;
    return (double)state / (double)m;
}

__global__ void __launch_bounds__(128) gen_zipf(ZipfGenParams p) {
    const uint64_t b = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= p.batches) return;
    const uint64_t first = p.batch * b, last = (b + 1 == p.batches) ? p.n : p.batch * (b + 1);
    int64_t state = p.base_seed + (int64_t)b;
    for (uint64_t i = first; i < last; ++i) {
        double sample;
        for (;;) {
            const double u1 = lcg_next(state);
            const double u2 = lcg_next(state);
            const double x = u1 * p.norm;
            const double inv = x <= 1.0 ? x : pow(x * p.sd + p.alpha, 1.0 / p.sd);
            sample = floor(inv + 1);
            const double p_target = pow(sample, -p.alpha);
            const double p_proposal = sample <= 1.0 ? 1.0 / p.norm : pow(inv, -p.alpha) / p.norm;
            if (u2 < p_target / (p_proposal * p.norm)) break;
        }
        p.out[i] = make_ulonglong2((uint64_t)((int64_t)sample + p.correction), i);
    }
}

// Test hook: raw hashes of a key array.
template <int HASH>
__global__ void hash_batch_kernel(const int64_t* __restrict__ keys, uint64_t n, HashParams hp,
                                  uint64_t* __restrict__ out) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = hash_key<HASH>((uint64_t)keys[i], hp);
}

}  // namespace phj
