/* include/phj.h -- C ABI of the B200-native hash-join engine (libphj_b200.so).
 *
 * This is the drop-in boundary for the one hot path of ragoragino/partitionedhashjoin ("phjoin"):
 * the count-only equi-join of two relations of 16-byte {int64 key, int64 payload} tuples, as a
 * no-partitioning join or a radix-partitioned join. The reference has no FFI today -- its joiners
 * are C++ templates called only from src/main.cpp:81-139 -- so each entry point below names the
 * reference interface it stands in for (file:line under /root/reference/). The C++ host mirror of
 * those interfaces lives in partitionedhashjoin_b200/host/ and is what INTEGRATION.md shows a
 * maintainer how to bind.
 *
 * Plain pointers and sizes only; no CUDA, torch or C++ types cross this boundary. Every function
 * returns 0 on success and a non-zero phj_status otherwise; phj_last_error() then describes the
 * failure (thread-local). There is NO CPU fallback: without a CUDA device every compute entry
 * point fails with PHJ_ERR_CUDA.
 */
#ifndef PHJ_H
#define PHJ_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PHJ_ABI_VERSION 5

/* Common::Tuple (src/Common/Table.hpp:20-25): alignas(16) {int64 id; int64 payload}. */
typedef struct {
    int64_t id;
    int64_t payload;
} phj_tuple;

/* Common::JoinedTuple (src/Common/Table.hpp:27-33). */
typedef struct {
    int64_t id;
    int64_t payloadA;
    int64_t payloadB;
} phj_joined_tuple;

/* Common::JoinAlgorithmType (src/Common/Configuration.hpp:12-15). */
enum {
    PHJ_ALGO_NO_PARTITIONING = 0,
    PHJ_ALGO_RADIX_PARTITIONING = 1,
    /* Multi-GPU exchange step (no counterpart in the single-process reference): phj_join only
     * splits the uploaded relations by the digit (hash >> shard_shift) % partitions, `partitions`
     * a power of two <= 256 (any value <= 256 when shard_shift is 0: the digit then is hash % partitions).
     * Either partitions = number of ranks (digit = owner rank; shard_shift beyond the local join's hash bits) and every rank then runs an ordinary RADIX_PARTITIONING
     * join on what it received; or partitions = ranks x local pass-1 digits with shard_shift =
     * the local join's radix_bits[1] (digit = owner rank : local pass-1 digit), in which case what
     * a rank receives is already pass-1 partitioned (phj_bind_device_partitioned). */
    PHJ_ALGO_SHARD_SPLIT = 2
};

/* Hasher behind Common::IHasher (src/Common/IHasher.hpp:6-11). The reference ships only XXH3
 * (src/Common/XXHasher.hpp:19-22); MURMUR3 and CITY are extensions named by the north star:
 *   MURMUR3 = low 64 bits of MurmurHash3_x64_128(&key, 8, (uint32)seed)
 *   CITY    = CityHash64WithSeed(&key, 8, seed), CityHash v1.0 short-key path (libc++ flavour) */
enum { PHJ_HASH_XXH3 = 0, PHJ_HASH_MURMUR3 = 1, PHJ_HASH_CITY = 2 };

enum {
    PHJ_OK = 0,
    PHJ_ERR_INVALID = 1, /* bad argument / configuration (std::invalid_argument in the reference) */
    PHJ_ERR_CUDA = 2,    /* CUDA runtime error, or no usable device */
    PHJ_ERR_STATE = 3,   /* call sequence error (e.g. join before upload) */
    PHJ_ERR_NOMEM = 4    /* device or pinned-host allocation failed */
};

/* Construction parameters: the union of what the reference passes to its two joiners --
 * NoPartitioning::Configuration / RadixClustering::Configuration (NumberOfPartitions,
 * src/RadixCluster/Configuration.hpp:6-9), the partition hasher and the table hasher
 * (src/main.cpp:213-217,267) -- plus the device. Zero-initialise, then set what you need. */
typedef struct {
    int32_t algo;        /* PHJ_ALGO_* */
    int32_t hash;        /* PHJ_HASH_* */
    uint64_t partitions; /* NumberOfPartitions; 0 = engine default. A power of two is split into
                            radix passes; any other value uses the reference's `hash % P` in one
                            pass (src/Common/XXHasher.hpp:21). Must be 0 for NO_PARTITIONING
                            (src/Arguments.hpp:12-17 rejects --partitions there). */
    uint32_t radix_bits[2]; /* optional: bits of pass 1 / pass 2 (most significant digit first);
                               {0,0} = derive from `partitions`. Sum must equal log2(partitions)
                               when both are given. */
    uint64_t hash_seed;  /* seed of the partitioning hasher (XXHasher::m_seed) */
    uint64_t table_seed; /* seed of the hash-table hasher; only observable through timing */
    int32_t device;      /* CUDA device ordinal */
    uint32_t flags;      /* PHJ_FLAG_* */
    uint64_t reserve_build; /* optional capacity hints (tuples); 0 = size on first upload */
    uint64_t reserve_probe;
    uint32_t shard_shift;   /* PHJ_ALGO_SHARD_SPLIT: first hash bit of the split digit */
    uint32_t split_ctas;    /* PHJ_ALGO_SHARD_SPLIT: cap on the CTAs of the split scatter (0 = one per
                               segment); an NVLink-bound split leaves the other SMs to a local join */
    uint32_t split_chunks;  /* PHJ_ALGO_SHARD_SPLIT: the probe relation is split in this many row chunks
                               (1..16, 0 = 1) that phj_shard_scatter sends one at a time; the build
                               relation travels with chunk 0 */
    uint32_t upload_chunks; /* phj_join_host: row chunks the probe relation is uploaded in, each joined
                               as soon as it has landed (1 = upload everything, then join; 0 = choose:
                               ~256 MB chunks when the probe relation is large, else 1; <= 32) */
    int32_t num_gpus;       /* 0 or 1 = one GPU. N > 1 (at most 16), RADIX_PARTITIONING: this process drives
                               GPUs device .. device + N - 1 -- phj_upload gives GPU g the rows [g n / N,
                               (g + 1) n / N) of both relations, phj_join runs the sharded join below (one host
                               thread per GPU, NCCL for sizes / barriers / the count, NVLink peer stores for the
                               tuples) and returns the global count. `partitions` then is GPUs x local
                               partitions (<= 256, a multiple of N, a power of two when N is one; 0 = 64, or the
                               largest multiple of N below it) and split_chunks the number of probe chunks whose
                               shuffle overlaps the local probe of the previous one (0 = 4). For N = 2, 4, 8, 16
                               the split digit is a bit field of the hash, for any other N it is hash % partitions.
                               NO_PARTITIONING with N > 1 needs no exchange: every GPU receives the
                               whole build relation and rows [g n / N, (g + 1) n / N) of the probe relation, builds
                               its own table and probes its share; phj_join returns the sum of the counts. */
    uint32_t reserved0;
} phj_config;

#define PHJ_FLAG_NO_TMA_STORE 0x2u /* scatter flush with st.global.v4 instead of TMA bulk stores */
#define PHJ_FLAG_SPLIT_REMOTE_ONLY 0x8u /* SHARD_SPLIT handle used only through phj_shard_scatter with
                                           both destinations given: no local output buffers */
#define PHJ_FLAG_SPLIT_LOCAL_TILES 0x40u /* SHARD_SPLIT: keep the HBM-tuned 4096-tuple tiles even when every
                                            digit has a destination (the split mostly writes local memory) */
#define PHJ_FLAG_CHAINED_TABLE 0x10u /* NO_PARTITIONING: bucket-chained global table (the reference's
                                        SeparateChainingHashTable, src/HashTables/SeparateChaining.hpp)
                                        instead of the open-addressing one (LinearProbing.hpp) */
#define PHJ_FLAG_FUSE_HIST2 0x4u   /* accepted for compatibility: pass 2 never reads a histogram of its own by default */
#define PHJ_FLAG_NO_FUSE_HIST2 0x20u /* pass 2 reads its own histogram. By default (two passes of <= 6 bits,
                                        power-of-two fan-out) ONE read of the input yields the histograms of
                                        both passes (radix_histogram_full: counts of (pass-1 digit, pass-2
                                        digit) pairs per pass-1 segment, summed into pass 2's counters); for
                                        other two-pass plans of <= 6 bits the pass-1 scatter counts them.
                                        Saves one read of both relations, DESIGN.md section 4 */
#define PHJ_FLAG_L2_TABLES 0x80u  /* RADIX_PARTITIONING: build + probe through per-partition tables in GLOBAL
                                     memory that stay hot in the 126 MB L2 (all CTAs walk the probe side in
                                     partition order) instead of shared-memory tables. Made for FEW, LARGE
                                     partitions: one radix pass (partitions = 32 .. 256) then needs no second
                                     partitioning pass at all -- 3.03-3.07 ms against 3.6-3.8 ms for two passes +
                                     shared memory at 10 M x 200 M. The multi-GPU path always joins this way. */
#define PHJ_FLAG_COOP_PROBE 0x200u /* global-table probes (NO_PARTITIONING, oversize fallback) by groups of four
                                      lanes per key (shuffle broadcast + one 8-byte slot per lane + ballot) instead
                                      of one thread per key reading the 32-byte bucket with one 256-bit load.
                                      Kept for the A/B in DESIGN.md section 4; the per-thread probe is faster. */
#define PHJ_FLAG_COUNT_UPFRONT 0x400u   /* sharded join: histogram the whole row shard before anything travels (default) */
#define PHJ_FLAG_COUNT_PIECEWISE 0x800u /* sharded join: count probe chunk c + 1 while chunk c travels. Hides 0.3 ms of
                                         * counting and gives it back through a slower shuffle: 5.81 against 5.87 ms at
                                         * 8 GPUs, slower at 2 (profiles/r02_multigpu.md). Opt-in, tested. */
#define PHJ_FLAG_HOT_DIGITS 0x1000u /* accepted for compatibility: heavy-hitter handling is the default */
#define PHJ_FLAG_NO_HOT_DIGITS 0x2000u /* sharded join: by default the split digits whose probe side alone outweighs a
                                     * quarter of one rank's fair share (Zipf heavy hitters; never the case for
                                     * uniform keys) keep their probe tuples where they are, and their build partition
                                     * is replicated to every rank: 5.84 against 10.07 ms at 8 GPUs and
                                     * Zipf 1.25. This flag sends every digit to its owner regardless. */
#define PHJ_FLAG_NO_HIST12 0x100u /* two-pass plans: do not take both passes' histograms from one read
                                     (radix_histogram_full); the pass-1 scatter counts for pass 2 instead */

/* What the reference reports through IHashJoinTimer (src/Common/Results.hpp:131-149) plus the
 * count it only logs (src/NoPartitioning/HashJoin.hpp:184, src/RadixCluster/HashJoin.hpp:320). */
typedef struct {
    uint64_t matches;      /* probe tuples with at least one build match */
    uint64_t partition_ns; /* device time: first partitioning kernel .. last scatter */
    uint64_t build_ns;     /* NPJ: table clear + build kernel. Radix: build share of the CTA with
                              the largest build+probe (RadixCluster/HashJoin.hpp:67-87,308-309) */
    uint64_t probe_ns;     /* NPJ: probe kernel only (the reference's figure also contains the
                              build, Results.hpp:202). Radix: probe share of that same CTA */
    uint64_t join_ns;      /* radix: device time of the fused build+probe kernel(s) */
    uint64_t total_ns;     /* device time, first kernel .. match count resident on the host */
    uint64_t h2d_ns;       /* device time of the uploads done by phj_join_host, else 0 */
    uint64_t hbm_bytes_alg; /* algorithmic HBM bytes of this join */
    uint32_t kernel_launches;
    uint32_t passes; /* radix passes actually run */
    uint64_t partitions; /* fan-out actually used */
    uint64_t fallback_partitions; /* partitions whose build side exceeded the shared-memory table
                                     and went through the global table instead */
    uint64_t h2d_bytes; /* bytes copied host -> device by this call (phj_join_host), else 0 */
    uint64_t d2h_bytes; /* bytes copied device -> host by this call (count, per-CTA phase times) */
    uint64_t joined_tuples;  /* phj_join_materialize: rows of the joined table, else 0 */
    uint64_t materialize_ns; /* phj_join_materialize: device time of the count + write kernels */
    uint64_t e2e_ns;         /* phj_join_host: device time, first upload .. match count on the host */
    uint32_t upload_chunks;  /* phj_join_host: probe chunks actually used (1 = not streamed), else 0 */
    uint32_t gpus;           /* GPUs that took part (sharded join), else 0 */
    uint64_t count_ns;       /* sharded join: histogram + scan + all-gather of the piece sizes */
    uint64_t shuffle_ns;     /* sharded join: first store into a peer's window .. last barrier (the local
                                probes of earlier chunks run inside this span) */
    uint64_t shuffle_bytes;  /* sharded join: bytes this rank (num_gpus > 1: all GPUs) sent to other GPUs */
} phj_result;

typedef struct phj_handle phj_handle;

/* ---- lifecycle ------------------------------------------------------------------------------ */

/* Stands in for constructing NoPartitioning::HashJoiner (src/NoPartitioning/HashJoin.hpp:43-52) /
 * RadixClustering::HashJoiner (src/RadixCluster/HashJoin.hpp:137-147) together with its hash-table
 * factory (src/main.cpp:216-217). */
int phj_create(const phj_config* config, phj_handle** out);
void phj_destroy(phj_handle* h);

/* Thread-local description of the last failure on this thread ("" if none). */
const char* phj_last_error(void);
uint32_t phj_abi_version(void);

/* ---- data ------------------------------------------------------------------------------------ */

/* Copy the build relation R (tableA) and probe relation S (tableB) to the device. Host pointers
 * are what a reference caller holds as &(*table)[0] of a Common::Table<Common::Tuple>
 * (src/Common/Table.hpp:35-57); the caller keeps ownership, the inputs are never modified.
 * Device arenas are (re)allocated here, i.e. outside phj_join's timed region, mirroring the
 * reference, which allocates its partition buffers before starting the timer
 * (src/RadixCluster/HashJoin.hpp:195-208). */
int phj_upload(phj_handle* h, const phj_tuple* build, size_t n_build, const phj_tuple* probe,
               size_t n_probe);

/* Same, for relations that already live in device memory of config.device (zero-copy: the engine
 * reads them in place and never writes them). */
int phj_bind_device(phj_handle* h, const void* d_build, size_t n_build, const void* d_probe,
                    size_t n_probe);

/* Same, for device-resident relations that are ALREADY partitioned by pass-1 digit: parent d lies at
 * [bounds[d], bounds[d + 1]), nparents + 1 host boundaries per relation, and holds the tuples whose
 * digit (hash >> radix_bits[1]) % parent_space equals first_parent + d. The join then starts at
 * pass 2 (or, for a one-pass plan, at build + probe) and produces nparents x 2^radix_bits[1]
 * partitions. parent_space = 0 means this handle's own 2^radix_bits[0]; a plain single-GPU use binds
 * all of them (first_parent 0, nparents = 2^radix_bits[0]). The multi-GPU path makes the NVLink
 * shuffle double as pass 1: the split digit space is ranks x local digits, and every rank owns a
 * contiguous digit range chosen to balance the tuple counts. Needs an explicit power-of-two
 * config.partitions; nparents <= 256. */
int phj_bind_device_partitioned(phj_handle* h, const void* d_build, size_t n_build, const void* d_probe,
                                size_t n_probe, const uint64_t* bounds_build, const uint64_t* bounds_probe,
                                uint32_t nparents, uint32_t first_parent, uint32_t parent_space);

/* Optional, after phj_bind_device_partitioned: the digit of every bound parent when they are not the
 * contiguous range first_parent .. (the multi-GPU path appends replicas of heavy-hitter digits owned
 * by other ranks). Only the handling of the reserved key value needs it; build partitions too large
 * for the shared-memory table are an error in this mode. */
int phj_set_parent_digits(phj_handle* h, const uint32_t* digits, uint32_t n);

/* ---- the join ---------------------------------------------------------------------------------
 * Stands in for HashJoiner::Run(tableA, tableB, timer) (src/NoPartitioning/HashJoin.hpp:54-74,
 * src/RadixCluster/HashJoin.hpp:190-241) on the relations given to phj_upload / phj_bind_device.
 * Synchronous; may be called repeatedly (the inputs stay resident). Not re-entrant per handle. */
int phj_join(phj_handle* h, phj_result* out);

/* The join WITH its result: fills the Table<JoinedTuple> that both reference Run()s return empty
 * (src/NoPartitioning/HashJoin.hpp:186, src/RadixCluster/HashJoin.hpp:226-227) -- one
 * phj_joined_tuple {id, payloadA (build), payloadB (probe)} per (probe tuple, equal-key build
 * tuple), i.e. LinearProbingHashTable::GetAll semantics (src/HashTables/LinearProbing.hpp:183-200).
 * Radix-partitioning handles only. out->matches is the count-only figure as in phj_join,
 * out->joined_tuples the number of rows; the rows live in a device buffer owned by the handle
 * (grown on demand inside this call), in no particular order, valid until the next call on the
 * handle. phj_read_joined copies rows [first, first + count) to the host; phj_device_joined
 * exposes the device buffer. */
int phj_join_materialize(phj_handle* h, phj_result* out);
int phj_read_joined(phj_handle* h, phj_joined_tuple* out, uint64_t first, uint64_t count);
int phj_device_joined(phj_handle* h, const void** d_joined, uint64_t* rows);

/* Upload + join in one call, the end-to-end path (host buffers in, count out) -- what the host
 * mirrors' Run(tableA, tableB) calls. With config.upload_chunks != 1 and a large probe relation the
 * call is STREAMED: the probe relation goes up in row chunks on a copy stream and every chunk is
 * joined against the whole build relation as soon as it has landed (counts are additive over a
 * partition of the probe relation), so only the last chunk's join is not hidden behind PCIe. Then
 * partition/build/probe/join/total_ns, hbm_bytes_alg and kernel_launches are sums over the chunks
 * (the build side is re-partitioned per chunk), h2d_ns is the device time of all uploads and
 * e2e_ns the device time of the whole call. Works from pageable memory too (uploads are issued by
 * a helper thread); pinned memory (phj_host_alloc) uploads faster. Both relations are resident
 * afterwards, as after phj_upload; phj_read_partitions needs a phj_join first. */
int phj_join_host(phj_handle* h, const phj_tuple* build, size_t n_build, const phj_tuple* probe,
                  size_t n_probe, phj_result* out);

/* ---- introspection / test hooks --------------------------------------------------------------- */

/* Device evaluation of the raw 64-bit hash (before `% cardinality`) of n host keys: pins the
 * __device__ hash functions bit-exactly against src/Common/XXHasher.hpp:19-22 from the host. */
int phj_hash_batch(int32_t hash, uint64_t seed, const int64_t* keys, size_t n, uint64_t* out,
                   int32_t device);
/* Host evaluation of the same function (the same source compiled for the host). */
uint64_t phj_hash_host(int32_t hash, uint64_t seed, int64_t key);

/* After a radix join: copy out the partitioned relation (`which` 0 = build, 1 = probe; n tuples)
 * and the partitions+1 partition boundaries -- the device analogue of partitionedTable +
 * PartitionsInfo (src/RadixCluster/HashJoin.hpp:16-33,195-198). Partition p = hash % P lies at
 * [bounds[p], bounds[p+1]) with the tuples in input order, i.e. bit-identical to the reference's
 * stable partitionTable (src/RadixCluster/HashJoin.hpp:394-412). Valid until the next upload/join. */
int phj_read_partitions(phj_handle* h, int32_t which, phj_tuple* out, uint64_t* bounds);

/* Device-resident view of the same: *d_data is the partitioned relation (n tuples), *d_bounds its
 * partitions+1 boundaries (uint64). Owned by the handle; valid until the next upload/join. Used by
 * the multi-GPU path to hand the SHARD_SPLIT output to the all-to-all without a host round trip. */
int phj_device_partitions(phj_handle* h, int32_t which, const void** d_data, const uint64_t** d_bounds,
                          size_t* n);

/* ---- multi-GPU exchange fused into the split (handles created with PHJ_ALGO_SHARD_SPLIT) ------
 * phj_shard_count: histogram + scan of the uploaded relations by split digit, once for all row
 * chunks: counts[(chunk * 2 + rel) * digits + digit] tuples of `chunk` carry `digit` (digits =
 * config.partitions, chunks = max(1, config.split_chunks)). phj_shard_scatter: the scatter of ONE
 * chunk, digit d's piece of relation `rel` written straight to dst[rel][d] + off[rel][d] tuples --
 * typically a peer GPU's receive window mapped with phj_shared_open, i.e. the partition shuffle
 * happens as NVLink stores from the scatter kernel (TMA bulk stores) instead of a separate
 * all-to-all. A null dst array keeps that relation local (split into the handle's own buffer); a
 * null ENTRY keeps that digit in that buffer too. Call order per join: phj_shard_count, exchange the counts, then per chunk:
 * phj_shard_scatter + a barrier across ranks. */
int phj_shard_count(phj_handle* h, uint64_t* counts);
int phj_shard_scatter(phj_handle* h, uint32_t chunk, void* const* dst_build, const uint64_t* off_build,
                      void* const* dst_probe, const uint64_t* off_probe, phj_result* out);

/* ---- the sharded radix join, one process per GPU ---------------------------------------
 * What RadixClustering::HashJoiner::Run (src/RadixCluster/HashJoin.hpp:190-241) becomes when R and S are
 * row-sharded over the GPUs of a node: the exchange is the one partitioning pass (every digit run is
 * stored straight into its owner's window over NVLink), the probe relation travels in chunks whose shuffle
 * overlaps the local probe of the previous chunk (the reference overlaps its two partition pipelines the
 * same way, :210-216), NCCL carries the piece sizes, the barriers and the count. All ranks call the same
 * functions in the same order (they are collectives). phj_nccl_unique_id: rank 0 obtains 128 opaque bytes
 * and hands them to the others over any transport (torch.distributed, MPI, a file). config as for
 * phj_create (algo RADIX_PARTITIONING; partitions = GPUs x local partitions as for num_gpus, 0 = choose;
 * split_chunks = probe chunks, 0 = 4; device = this rank's GPU). phj_dist_join returns the GLOBAL count on
 * every rank. The same join inside ONE process: phj_config.num_gpus. */
typedef struct phj_dist phj_dist;
int phj_nccl_unique_id(unsigned char* id128);
int phj_dist_create(const phj_config* config, int32_t rank, int32_t world, const unsigned char* id128,
                    phj_dist** out);
void phj_dist_destroy(phj_dist* d);
int phj_dist_upload(phj_dist* d, const phj_tuple* build, size_t n_build, const phj_tuple* probe, size_t n_probe);
int phj_dist_bind_device(phj_dist* d, const void* d_build, size_t n_build, const void* d_probe, size_t n_probe);
int phj_dist_join(phj_dist* d, phj_result* out);
int phj_dist_kernel_times(phj_dist* d, const char** names, uint64_t* ns, uint32_t cap);
/* Which kernels of the following joins get CUDA events around them, as phj_kernel_timing: null = none, "" = all,
 * else those whose name contains `filter`. */
int phj_dist_kernel_timing(phj_dist* d, const char* filter);
/* Same launches with their begin / end device times relative to the join's first event (needs
 * PHJ_KERNEL_TIMES=1): shows which kernels of the two streams ran side by side. */
int phj_dist_kernel_trace(phj_dist* d, const char** names, uint64_t* begin_ns, uint64_t* end_ns, uint32_t cap);

/* NVLink reference measured in place: every rank copies `bytes` of its probe window into the next rank's
 * window with a plain cudaMemcpyAsync (one per repeat), all ranks at once -- the per-direction bandwidth the
 * shuffle is compared with (bench.py: shuffle.nvlink_peak_GBps). Collective; overwrites the windows. */
int phj_dist_measure_peer_copy(phj_dist* d, uint64_t bytes, uint32_t repeats, uint64_t* ns_per_copy);

/* How the last join laid this rank's data out, and test read-back of its windows: `which` 0 = build
 * (bounds: partitions_here + 1), 1 = probe (bounds: chunks x (partitions_here + 1), absolute positions:
 * chunk c's tuples of local partition l lie at [bounds[c][l], bounds[c][l + 1]), ordered by source rank,
 * then input order). Partitions [local_partitions, partitions_here) are the hot digits of other owners in digit
 * order: the whole build partition (copied from its owner), this rank's own probe tuples. The window's used
 * prefix (bounds' last entry) is copied to `out` when it is given. */
typedef struct {
    uint32_t world, rank;
    uint32_t digits;            /* split digits = world x local_partitions; owner = digit / local_partitions; digit =
                                   hash & (digits - 1), or hash % digits when digits is not a power of two */
    uint32_t local_partitions;
    uint32_t chunks;
    uint32_t region_buckets;    /* 32-byte buckets per local partition's table */
    uint64_t window_tuples[2];  /* capacity of this rank's build / probe window */
    uint32_t resizes;           /* host-synchronous sizing passes that re-allocated a window so far */
    uint32_t partitions_here;   /* partitions in this rank's windows: local_partitions + the hot digits of other owners */
    uint64_t sent_remote_bytes;
    uint32_t hot_count;         /* PHJ_FLAG_HOT_DIGITS: split digits kept local / replicated, ascending */
    uint32_t hot_digits[32];
    uint32_t reserved;
} phj_dist_layout;
int phj_dist_info(phj_dist* d, phj_dist_layout* out);
int phj_dist_read_window(phj_dist* d, int32_t which, phj_tuple* out, uint64_t cap_tuples, uint64_t* bounds);

/* Inside ONE process a window of another GPU needs no IPC handle: enable peer access from `device` to `peer`
 * once and use the pointer phj_shared_alloc returned (what phj_config.num_gpus does internally). */
int phj_enable_peer_access(int32_t device, int32_t peer);

/* Device memory that other processes on the node can map (CUDA IPC): the receive buffers of the
 * fused shuffle. `ipc_handle` is 64 opaque bytes to hand to the peers (any transport). */
int phj_shared_alloc(int32_t device, size_t bytes, void** d_ptr, unsigned char* ipc_handle);
int phj_shared_open(int32_t device, const unsigned char* ipc_handle, void** d_ptr);
int phj_shared_close(int32_t device, void* d_ptr);
int phj_shared_free(int32_t device, void* d_ptr);

/* Device-side input generation for scale runs (the 160 M x 3.2 B configuration): the reference's
 * Sequential / Zipf generators (src/DataGenerator/Sequential.cpp:20-25, Zipf.cpp:14-56,80-92) with
 * the batch / seed scheme of phj_fill_zipf, one GPU thread per batch, written to device memory
 * (phj_shared_alloc or any device pointer), ready for phj_bind_device. CUDA's pow() is not glibc's:
 * a sample may differ from phj_fill_zipf's where an intermediate lands within an ulp of an integer,
 * so parity runs use the host generators and upload. */
int phj_device_fill_sequential(int32_t device, void* d_out, size_t n, int64_t start);
int phj_device_fill_zipf(int32_t device, void* d_out, size_t n, double alpha, int64_t range_first,
                         int64_t range_second, int64_t base_seed, size_t batches);

/* Plain synchronous copies between host memory and memory from phj_shared_alloc (input arenas of
 * the multi-GPU path, test read-back of the windows). */
int phj_memcpy_h2d(int32_t device, void* d_dst, const void* h_src, size_t bytes);
int phj_memcpy_d2h(int32_t device, void* h_dst, const void* d_src, size_t bytes);
int phj_memcpy_d2d(int32_t device, void* d_dst, const void* d_src, size_t bytes); /* incl. mapped peer memory */

/* Which kernels of the following joins get CUDA events around them: null = none, "" = all, else those whose
 * name contains `filter` (what PHJ_KERNEL_TIMES = 1 / a name selects at phj_create). */
int phj_kernel_timing(phj_handle* h, const char* filter);

/* Per-kernel device times of the last phj_join: up to `cap` entries; returns the number written.
 * names[i] points to a static string. */
int phj_kernel_times(phj_handle* h, const char** names, uint64_t* ns, uint32_t cap);

/* Device facts the host side prints next to results (SM count, L2, HBM, clocks). */
typedef struct {
    char name[128];
    int32_t sm_count;
    int32_t cc_major, cc_minor;
    uint64_t global_mem_bytes;
    uint64_t l2_bytes;
    uint64_t smem_per_block_optin;
    int32_t sm_clock_khz, mem_clock_khz, mem_bus_bits;
} phj_device_info;
int phj_get_device_info(int32_t device, phj_device_info* out);
int phj_device_count(void);

/* ---- data generators (host side) ---------------------------------------------------------------
 * DataGenerator::Sequential / DataGenerator::Zipf (src/DataGenerator/Sequential.cpp:6-40,
 * src/DataGenerator/Zipf.cpp:14-108) and the Park-Miller LCG (src/Common/Random.cpp:9-30). They
 * produce the join's INPUT on the host exactly as the reference does (device pow() is not
 * bit-identical to glibc's), with the seeding made explicit: the table is cut into `batches`
 * equal slices (the last takes the remainder) and slice b draws from an LCG seeded
 * base_seed + b. `threads` <= 0 uses all hardware threads. */
int phj_fill_sequential(phj_tuple* out, size_t n, int64_t start, int32_t threads);
int phj_fill_zipf(phj_tuple* out, size_t n, double alpha, int64_t range_first, int64_t range_second,
                  int64_t base_seed, size_t batches, int32_t threads);

/* Pinned host memory for relations that are uploaded repeatedly (phj_join_host). */
int phj_host_alloc(void** out, size_t bytes);
int phj_host_free(void* p);

#ifdef __cplusplus
}
#endif
#endif /* PHJ_H */
