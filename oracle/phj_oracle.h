/* oracle/phj_oracle.h -- TEST INFRASTRUCTURE ONLY.
 *
 * Plain-C restatement of the reference's (ragoragino/partitionedhashjoin) hash-join hot path, used
 * by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg as the CHECKER for the CUDA
 * path. The product (partitionedhashjoin_b200/) never includes, links or calls this.
 *
 * Parity status: PINNED. Every function below is checked by tests/test_oracle.py against
 *   (a) the golden vectors of SURVEY.md Appendix A (committed under tests/golden/), and
 *   (b) the UNMODIFIED reference compiled into oracle/_ref/libphj_ref.so (oracle/ref_harness.cpp),
 * except the Murmur3 / City hashers, which do not exist in the reference (parity unpinned by the
 * reference; pinned instead to the canonical sources named at their definitions).
 *
 * All citations are file:line under /root/reference/.
 */
#ifndef PHJ_ORACLE_H
#define PHJ_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* src/Common/Table.hpp:20-25 */
typedef struct {
    int64_t id;
    int64_t payload;
} phjo_tuple;

/* ---- hashing (src/Common/XXHasher.hpp:19-22; arithmetic from xxHash >= 0.8.0) ---- */
uint64_t phjo_xxh3_64(int64_t key, uint64_t seed);
uint64_t phjo_murmur3_64(int64_t key, uint64_t seed); /* extension, not in the reference */
uint64_t phjo_city_64(int64_t key, uint64_t seed);    /* extension, not in the reference */
/* hash_id: 0 xxh3, 1 murmur3, 2 city. Returns hash % cardinality like IHasher::Hash. */
uint64_t phjo_hash(int hash_id, int64_t key, uint64_t seed, uint64_t cardinality);
uint64_t phjo_hash_raw(int hash_id, int64_t key, uint64_t seed);
void phjo_hash_batch(int hash_id, uint64_t seed, const int64_t* keys, size_t n, uint64_t* out);

/* ---- generators (src/Common/Random.cpp:9-30, src/DataGenerator/{Zipf,Sequential}.cpp) ---- */
double phjo_lcg_next(long* state);
/* returns 0 on success, 1 if alpha < 0.01 (the reference throws, Zipf.cpp:18-20) */
int phjo_zipf_generate(double alpha, uint64_t cardinality, long* state, uint64_t* sample);
void phjo_fill_sequential(phjo_tuple* out, size_t n, int64_t start);
int phjo_fill_zipf_seeded(phjo_tuple* out, size_t n, double alpha, int64_t range_first,
                          int64_t range_second, long base_seed, size_t batches);

/* ---- hash tables (src/HashTables/{LinearProbing,SeparateChaining}.hpp) ---- */
typedef struct phjo_table phjo_table;
/* kind: 0 linear probing, 1 separate chaining. size_ratio <= 0 selects the reference default
 * (1.25 / 0.25). Returns NULL for number_of_objects == 0 (the reference throws). */
phjo_table* phjo_table_new(int kind, double size_ratio, int hash_id, uint64_t seed,
                           size_t number_of_objects);
void phjo_table_free(phjo_table* t);
uint64_t phjo_table_buckets(const phjo_table* t);
/* returns 0, or 1 when the separate-chaining overflow allocator is exhausted (the reference
 * throws "BucketAllocator exceeded its limit.", SeparateChaining.hpp:114-118) */
int phjo_table_insert(phjo_table* t, int64_t key, const phjo_tuple* value);
int phjo_table_exists(const phjo_table* t, int64_t key);
const phjo_tuple* phjo_table_get(const phjo_table* t, int64_t key);
size_t phjo_table_get_all(const phjo_table* t, int64_t key, const phjo_tuple** out, size_t cap);

/* ---- joins ---- */
typedef struct {
    uint64_t matches; /* probe tuples with >= 1 build match (Get != nullptr) */
    uint64_t partition_ns, build_ns, probe_ns;
} phjo_result;

/* src/NoPartitioning/HashJoin.hpp:54-187. Returns 1 when nR == 0 (table ctor throws). */
int phjo_join_no_partitioning(const phjo_tuple* R, size_t nR, const phjo_tuple* S, size_t nS,
                              int table_kind, int hash_id, uint64_t seed_table, phjo_result* out);

/* src/Common/Table.hpp:27-33 */
typedef struct {
    int64_t id;
    int64_t payloadA;
    int64_t payloadB;
} phjo_joined;

/* The joined table the reference declares but never fills: one row per (probe tuple, equal-key
 * build tuple), via GetAll (src/HashTables/LinearProbing.hpp:183-200). Returns the row count;
 * writes at most cap rows. UINT64_MAX if the chaining allocator overflowed. */
uint64_t phjo_join_materialize(const phjo_tuple* R, size_t nR, const phjo_tuple* S, size_t nS,
                               int table_kind, int hash_id, uint64_t seed_table, phjo_joined* out,
                               uint64_t cap);

/* src/RadixCluster/HashJoin.hpp:149-188 -- worker count / batch sizes for both relations. */
void phjo_partitioning_configuration(size_t sizeA, size_t sizeB, size_t pool_workers,
                                     size_t min_batch, size_t* workers, size_t* batchA,
                                     size_t* batchB);

/* src/RadixCluster/HashJoin.hpp:333-440 -- histogram / prefix sum / stable scatter of ONE relation
 * with `workers` workers of `batch` tuples (the last takes the remainder, :420-422).
 * out: n tuples; bounds: 2*P entries, (first, second) per partition (:18-25).
 * prefix (optional, may be NULL): workers*P entries = the exclusive per-partition prefix over
 * workers as left by createPrefixSumTable (:363-390), worker-major like PrefixSumTable (:46-47). */
void phjo_radix_partition(const phjo_tuple* in, size_t n, size_t P, int hash_id, uint64_t seed,
                          size_t workers, size_t batch, phjo_tuple* out, uint64_t* bounds,
                          uint64_t* prefix);

/* src/RadixCluster/HashJoin.hpp:190-331 -- whole radix join with `pool_workers` workers. */
int phjo_join_radix(const phjo_tuple* R, size_t nR, const phjo_tuple* S, size_t nS, size_t P,
                    size_t pool_workers, int table_kind, int hash_id, uint64_t seed_partition,
                    uint64_t seed_table, phjo_result* out);

/* Independent O(n log n) cross-check of the semi-join count (sort + binary search; no hashing). */
uint64_t phjo_count_by_sort(const phjo_tuple* R, size_t nR, const phjo_tuple* S, size_t nS);

#ifdef __cplusplus
}
#endif
#endif
