/* oracle/phj_oracle.c -- TEST INFRASTRUCTURE ONLY (see phj_oracle.h).
 *
 * CPU restatement, in plain C, of the algorithm of ragoragino/partitionedhashjoin's join hot path.
 * It follows the reference's control flow function by function (citations: file:line under
 * /root/reference/) but shares no code with it, and runs the reference's W "workers"
 * sequentially (the results of every loop restated here are independent of thread interleaving;
 * where the reference's result depends on it -- which duplicate Get() returns after a concurrent
 * build -- the tests only compare interleaving-independent quantities).
 *
 * Parity: PINNED by tests/test_oracle.py against tests/golden/ and oracle/_ref/libphj_ref.so.
 */
#define _POSIX_C_SOURCE 199309L
#include "phj_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

/* ================================================================================================
 * Hashing
 * ============================================================================================== */

static inline uint64_t rotl64(uint64_t x, int r) { return (x << r) | (x >> (64 - r)); }
static inline uint64_t rotr64(uint64_t x, int r) { return (x >> r) | (x << (64 - r)); }
static inline uint32_t bswap32(uint32_t x) {
    return (x >> 24) | ((x >> 8) & 0xff00u) | ((x << 8) & 0xff0000u) | (x << 24);
}

/* XXH3_64bits_withSeed(&key, 8, seed), the only call XXHasher makes (src/Common/XXHasher.hpp:20).
 * The arithmetic lives in third-party xxHash (find_package(xxHash 0.7), CMakeLists.txt:14; not
 * vendored). Restated from the published, frozen (>= 0.8.0) algorithm: XXH3_64bits_withSeed ->
 * XXH3_len_0to16_64b -> XXH3_len_4to8_64b -> XXH3_rrmxmx, with kSecret bytes 8..23. */
uint64_t phjo_xxh3_64(int64_t key, uint64_t seed) {
    const uint64_t secret8 = 0x1cad21f72c81017cULL;  /* LE64(kSecret + 8)  */
    const uint64_t secret16 = 0xdb979083e96dd4deULL; /* LE64(kSecret + 16) */
    const uint64_t prime_mx2 = 0x9FB21C651E98DF25ULL;
    uint64_t k = (uint64_t)key;
    seed ^= (uint64_t)bswap32((uint32_t)seed) << 32;
    uint64_t bitflip = (secret8 ^ secret16) - seed;
    uint64_t input64 = (k >> 32) + (k << 32); /* input2 + (input1 << 32) */
    uint64_t h = input64 ^ bitflip;
    h ^= rotl64(h, 49) ^ rotl64(h, 24);
    h *= prime_mx2;
    h ^= (h >> 35) + 8; /* len == 8 */
    h *= prime_mx2;
    h ^= h >> 28;
    return h;
}

/* Extension (no counterpart in the reference): low 64 bits (h1) of Appleby's canonical
 * MurmurHash3_x64_128(&key, 8, (uint32_t)seed); pinned in tests against the copy of the canonical
 * source that scikit-learn ships (sklearn/utils/src/MurmurHash3.cpp). */
static inline uint64_t fmix64(uint64_t k) {
    k ^= k >> 33;
    k *= 0xff51afd7ed558ccdULL;
    k ^= k >> 33;
    k *= 0xc4ceb9fe1a85ec53ULL;
    k ^= k >> 33;
    return k;
}
uint64_t phjo_murmur3_64(int64_t key, uint64_t seed) {
    const uint64_t c1 = 0x87c37b91114253d5ULL, c2 = 0x4cf5ad432745937fULL;
    uint64_t h1 = (uint32_t)seed, h2 = (uint32_t)seed;
    uint64_t k1 = (uint64_t)key; /* tail, len & 15 == 8 */
    k1 *= c1;
    k1 = rotl64(k1, 31);
    k1 *= c2;
    h1 ^= k1;
    h1 ^= 8;
    h2 ^= 8;
    h1 += h2;
    h2 += h1;
    h1 = fmix64(h1);
    h2 = fmix64(h2);
    h1 += h2;
    return h1;
}

/* Extension (no counterpart in the reference): CityHash64WithSeed of the 8 key bytes using the
 * CityHash v1.0 short-key path that libc++ / libcudacxx ship (len 4..8 ->
 * HashLen16(len + (lo32 << 3), hi32)), then WithSeed: HashLen16(h - k2, seed). Pinned in tests
 * against cuda::std::__murmur2_or_cityhash<size_t, 64> from the CCCL headers in this image. */
static inline uint64_t city_hash_len_16(uint64_t u, uint64_t v) {
    const uint64_t mul = 0x9ddfea08eb382d69ULL;
    uint64_t a = (u ^ v) * mul;
    a ^= a >> 47;
    uint64_t b = (v ^ a) * mul;
    b ^= b >> 47;
    b *= mul;
    return b;
}
uint64_t phjo_city_64(int64_t key, uint64_t seed) {
    const uint64_t k2 = 0x9ae16a3b2f90404fULL;
    uint64_t k = (uint64_t)key;
    uint64_t lo = (uint32_t)k, hi = (uint32_t)(k >> 32);
    uint64_t h = city_hash_len_16(8 + (lo << 3), hi);
    return city_hash_len_16(h - k2, seed);
}

uint64_t phjo_hash_raw(int hash_id, int64_t key, uint64_t seed) {
    switch (hash_id) {
        case 1:
            return phjo_murmur3_64(key, seed);
        case 2:
            return phjo_city_64(key, seed);
        default:
            return phjo_xxh3_64(key, seed);
    }
}

/* IHasher::Hash: hash % cardinality (src/Common/XXHasher.hpp:21). */
uint64_t phjo_hash(int hash_id, int64_t key, uint64_t seed, uint64_t cardinality) {
    return phjo_hash_raw(hash_id, key, seed) % cardinality;
}

void phjo_hash_batch(int hash_id, uint64_t seed, const int64_t* keys, size_t n, uint64_t* out) {
    for (size_t i = 0; i != n; ++i) out[i] = phjo_hash_raw(hash_id, keys[i], seed);
}

/* ================================================================================================
 * Generators
 * ============================================================================================== */

/* MultiplicativeLCGRandomNumberGenerator::Next (src/Common/Random.cpp:9-30): Park-Miller minimal
 * standard via Schrage's decomposition, state kept as `long`, returns state / m as a double. */
double phjo_lcg_next(long* state) {
    const long a = 16807, m = 2147483647, q = 127773, r = 2836;
    long x_div_q = *state / q;
    long x_mod_q = *state % q;
    long x_new = a * x_mod_q - r * x_div_q;
    *state = x_new > 0 ? x_new : x_new + m;
    return (double)*state / (double)m;
}

/* Zipf::generate (src/DataGenerator/Zipf.cpp:14-56): rejection-inversion sampling. */
int phjo_zipf_generate(double alpha, uint64_t cardinality, long* state, uint64_t* sample_out) {
    const double error_differential = 0.01;
    if (alpha < 0.01) return 1; /* :18-20 throws */

    double skew_differential = 1.001 - alpha; /* :22 */
    double diff = 1.0 - alpha;
    if (fabs(diff) < error_differential) { /* :23-26 */
        skew_differential = error_differential * ((diff < 0) ? 1 : -1);
        alpha = 1.0 - skew_differential;
    }
    double norm = (pow((double)cardinality, skew_differential) - alpha) / skew_differential; /* :28 */

    for (;;) {
        double u1 = phjo_lcg_next(state); /* :32-33 */
        double u2 = phjo_lcg_next(state);
        double inv; /* :37-43 */
        if (u1 * norm <= 1.0) {
            inv = u1 * norm;
        } else {
            inv = pow((u1 * norm) * skew_differential + alpha, 1.0 / skew_differential);
        }
        double sample = floor(inv + 1);                                      /* :45 */
        double density_original = pow(sample, -alpha);                       /* :46 */
        double density_sampling =                                            /* :47-49 */
            sample <= 1.0 ? 1.0 / norm : pow(inv, -alpha) / norm;
        double ratio = density_original / (density_sampling * norm);         /* :50 */
        if (u2 < ratio) {                                                    /* :52-54 */
            *sample_out = (uint64_t)sample;
            return 0;
        }
    }
}

/* Sequential::FillTable's per-element rule (src/DataGenerator/Sequential.cpp:20-25): batches are
 * contiguous and start at parameters.start + batch offset (:37), so id = start + i overall. */
void phjo_fill_sequential(phjo_tuple* out, size_t n, int64_t start) {
    for (size_t i = 0; i != n; ++i) {
        out[i].id = start + (int64_t)i;
        out[i].payload = (int64_t)i;
    }
}

/* Zipf::FillTable's per-element rule (src/DataGenerator/Zipf.cpp:80-92) with the batch split and
 * seeding fixed the same way oracle/ref_harness.cpp:phjref_fill_zipf_seeded fixes them (the
 * reference seeds from std::random_device, :86): `batches` equal slices, the last takes the
 * remainder (:99-101), slice b uses an LCG seeded base_seed + b. */
int phjo_fill_zipf_seeded(phjo_tuple* out, size_t n, double alpha, int64_t range_first,
                          int64_t range_second, long base_seed, size_t batches) {
    if (range_first >= range_second || batches == 0) return 1; /* :61-67 */
    size_t batch = n / batches;
    int64_t cardinality = range_second - range_first + 1; /* :81 */
    int64_t correction = range_first - 1;                 /* :82-83 */
    for (size_t b = 0; b != batches; ++b) {
        size_t start = batch * b, end = (b + 1 == batches) ? n : batch * (b + 1);
        long state = base_seed + (long)b;
        for (size_t i = start; i != end; ++i) {
            uint64_t s;
            if (phjo_zipf_generate(alpha, (uint64_t)cardinality, &state, &s)) return 1;
            out[i].id = (int64_t)s + correction; /* :90 */
            out[i].payload = (int64_t)i;         /* :91 */
        }
    }
    return 0;
}

/* ================================================================================================
 * Hash tables
 * ============================================================================================== */

#define PHJO_BUCKET_SLOTS 3 /* TupleSize = 3, src/main.cpp:211 */

/* internal::LinearProbing::Bucket / internal::SeparateChaining::Bucket
 * (src/HashTables/LinearProbing.hpp:22-83, SeparateChaining.hpp:22-101): an append-only array of
 * 3 (key, value*) slots plus m_freePosition; the chaining bucket adds a next pointer. */
typedef struct phjo_bucket {
    struct phjo_bucket* next;
    int8_t free_position;
    int64_t keys[PHJO_BUCKET_SLOTS];
    const phjo_tuple* values[PHJO_BUCKET_SLOTS];
} phjo_bucket;

struct phjo_table {
    int kind;
    int hash_id;
    uint64_t seed;
    uint64_t number_of_buckets;
    phjo_bucket* buckets;      /* LP: the table; SC: m_firstBuckets */
    phjo_bucket** bucket_ptrs; /* SC only: m_bucketPtrs (chain heads) */
    phjo_bucket* overflow;     /* SC only: BucketAllocator storage */
    size_t overflow_cap, overflow_used;
};

static int bucket_insert(phjo_bucket* b, int64_t key, const phjo_tuple* v) {
    if (b->free_position == PHJO_BUCKET_SLOTS) return 0; /* LinearProbing.hpp:33-36 */
    b->keys[b->free_position] = key;
    b->values[b->free_position] = v;
    b->free_position++;
    return 1;
}
static const phjo_tuple* bucket_get(const phjo_bucket* b, int64_t key) {
    for (int i = 0; i != b->free_position; ++i) /* LinearProbing.hpp:60-70 */
        if (b->keys[i] == key) return b->values[i];
    return NULL;
}

/* getNumberOfBuckets: ceil(ratio * n) (src/HashTables/LinearProbing.cpp:7-12,
 * src/HashTables/SeparateChaining.cpp:7-13); default ratios LinearProbing.hpp:17 (1.25),
 * SeparateChaining.hpp:17 (0.25). */
phjo_table* phjo_table_new(int kind, double size_ratio, int hash_id, uint64_t seed,
                           size_t number_of_objects) {
    if (number_of_objects == 0) return NULL; /* LinearProbing.hpp:106-110 throws */
    if (size_ratio <= 0) size_ratio = kind == 0 ? 1.25 : 0.25;
    phjo_table* t = (phjo_table*)calloc(1, sizeof(*t));
    t->kind = kind;
    t->hash_id = hash_id;
    t->seed = seed;
    t->number_of_buckets = (uint64_t)ceil(size_ratio * (double)number_of_objects);
    t->buckets = (phjo_bucket*)calloc(t->number_of_buckets, sizeof(phjo_bucket));
    if (kind == 1) {
        t->bucket_ptrs = (phjo_bucket**)calloc(t->number_of_buckets, sizeof(phjo_bucket*));
        /* SeparateChaining.hpp:165-168: ceil(n / BucketSize) overflow buckets */
        t->overflow_cap = (size_t)ceil((double)number_of_objects / (double)PHJO_BUCKET_SLOTS);
        t->overflow = (phjo_bucket*)calloc(t->overflow_cap, sizeof(phjo_bucket));
    }
    return t;
}

void phjo_table_free(phjo_table* t) {
    if (!t) return;
    free(t->buckets);
    free(t->bucket_ptrs);
    free(t->overflow);
    free(t);
}

uint64_t phjo_table_buckets(const phjo_table* t) { return t->number_of_buckets; }

int phjo_table_insert(phjo_table* t, int64_t key, const phjo_tuple* value) {
    uint64_t h = phjo_hash(t->hash_id, key, t->seed, t->number_of_buckets);
    if (t->kind == 0) {
        /* LinearProbingHashTable::Insert (LinearProbing.hpp:114-134); the per-bucket spin latch
         * (:120-126) is a no-op single-threaded. */
        for (;;) {
            if (bucket_insert(&t->buckets[h], key, value)) return 0;
            h = (++h == t->number_of_buckets) ? 0 : h;
        }
    }
    /* SeparateChainingHashTable::Insert (SeparateChaining.hpp:172-213) */
    if (t->bucket_ptrs[h] == NULL) {
        t->bucket_ptrs[h] = &t->buckets[h];
        bucket_insert(t->bucket_ptrs[h], key, value);
    } else if (!bucket_insert(t->bucket_ptrs[h], key, value)) {
        if (t->overflow_used >= t->overflow_cap) return 1; /* BucketAllocator::New throws :114-118 */
        phjo_bucket* fresh = &t->overflow[t->overflow_used++];
        fresh->next = t->bucket_ptrs[h]; /* pushed at the chain head, :199-201 */
        t->bucket_ptrs[h] = fresh;
        bucket_insert(fresh, key, value);
    }
    return 0;
}

const phjo_tuple* phjo_table_get(const phjo_table* t, int64_t key) {
    uint64_t h = phjo_hash(t->hash_id, key, t->seed, t->number_of_buckets);
    if (t->kind == 0) {
        /* LinearProbingHashTable::Get (LinearProbing.hpp:160-180): scan the bucket; stop at the
         * first bucket that is not full. */
        for (;;) {
            const phjo_bucket* b = &t->buckets[h];
            const phjo_tuple* v = bucket_get(b, key);
            if (v) return v;
            if (b->free_position != PHJO_BUCKET_SLOTS) return NULL;
            h = (++h == t->number_of_buckets) ? 0 : h;
        }
    }
    /* SeparateChainingHashTable::Get (SeparateChaining.hpp:238-258) */
    for (const phjo_bucket* b = t->bucket_ptrs[h]; b != NULL; b = b->next) {
        const phjo_tuple* v = bucket_get(b, key);
        if (v) return v;
    }
    return NULL;
}

/* Exists (LinearProbing.hpp:137-157, SeparateChaining.hpp:216-235) has Get's control flow. */
int phjo_table_exists(const phjo_table* t, int64_t key) { return phjo_table_get(t, key) != NULL; }

size_t phjo_table_get_all(const phjo_table* t, int64_t key, const phjo_tuple** out, size_t cap) {
    uint64_t h = phjo_hash(t->hash_id, key, t->seed, t->number_of_buckets);
    size_t n = 0;
    if (t->kind == 0) {
        /* LinearProbingHashTable::GetAll (LinearProbing.hpp:183-200) */
        for (;;) {
            const phjo_bucket* b = &t->buckets[h];
            for (int i = 0; i != b->free_position; ++i)
                if (b->keys[i] == key) {
                    if (out && n < cap) out[n] = b->values[i];
                    ++n;
                }
            if (b->free_position != PHJO_BUCKET_SLOTS) return n;
            h = (++h == t->number_of_buckets) ? 0 : h;
        }
    }
    /* SeparateChaining Bucket::GetAll walks the chain (SeparateChaining.hpp:73-93, :261-269) */
    for (const phjo_bucket* b = t->bucket_ptrs[h]; b != NULL; b = b->next)
        for (int i = 0; i != b->free_position; ++i)
            if (b->keys[i] == key) {
                if (out && n < cap) out[n] = b->values[i];
                ++n;
            }
    return n;
}

/* ================================================================================================
 * Joins
 * ============================================================================================== */

static uint64_t now_ns(void) {
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (uint64_t)ts.tv_sec * 1000000000ull + (uint64_t)ts.tv_nsec;
}

/* NoPartitioning::HashJoiner::Run / Build / Probe (src/NoPartitioning/HashJoin.hpp:54-74, 76-126,
 * 128-187): one table sized for |R| (:82), every R tuple inserted with value = &tuple (:93-98),
 * then counter++ for every S tuple whose Get is non-null (:144-153). The worker split only
 * changes who increments which partial counter (:155). */
int phjo_join_no_partitioning(const phjo_tuple* R, size_t nR, const phjo_tuple* S, size_t nS,
                              int table_kind, int hash_id, uint64_t seed_table, phjo_result* out) {
    memset(out, 0, sizeof(*out));
    uint64_t t0 = now_ns();
    phjo_table* table = phjo_table_new(table_kind, 0, hash_id, seed_table, nR);
    if (!table) return 1;
    for (size_t i = 0; i != nR; ++i)
        if (phjo_table_insert(table, R[i].id, &R[i])) {
            phjo_table_free(table);
            return 2;
        }
    uint64_t t1 = now_ns();
    uint64_t counter = 0;
    for (size_t i = 0; i != nS; ++i)
        if (phjo_table_get(table, S[i].id) != NULL) counter++;
    uint64_t t2 = now_ns();
    phjo_table_free(table);
    out->matches = counter;
    out->build_ns = t1 - t0;
    out->probe_ns = t2 - t1;
    return 0;
}

/* The joined table both Run()s declare but leave empty (src/NoPartitioning/HashJoin.hpp:186,
 * src/RadixCluster/HashJoin.hpp:226-227): Table<JoinedTuple>{id, payloadA, payloadB}
 * (src/Common/Table.hpp:27-33), filled the way the probe loop would with GetAll
 * (src/HashTables/LinearProbing.hpp:183-200): one row per (probe tuple, equal-key build tuple).
 * The reference itself never executes this loop, so this is the definition the CUDA path is held
 * to, not a restatement of executed reference code. Rows are emitted in probe order; rows of one
 * probe tuple follow the table's bucket order. Returns the number of joined rows (even beyond cap). */
uint64_t phjo_join_materialize(const phjo_tuple* R, size_t nR, const phjo_tuple* S, size_t nS,
                               int table_kind, int hash_id, uint64_t seed_table, phjo_joined* out,
                               uint64_t cap) {
    if (nR == 0) return 0;
    phjo_table* table = phjo_table_new(table_kind, 0, hash_id, seed_table, nR);
    if (!table) return 0;
    for (size_t i = 0; i != nR; ++i)
        if (phjo_table_insert(table, R[i].id, &R[i])) {
            phjo_table_free(table);
            return UINT64_MAX;
        }
    size_t scratch_cap = 64;
    const phjo_tuple** scratch = (const phjo_tuple**)malloc(scratch_cap * sizeof(*scratch));
    uint64_t n = 0;
    for (size_t i = 0; i != nS; ++i) {
        size_t k = phjo_table_get_all(table, S[i].id, scratch, scratch_cap);
        if (k > scratch_cap) {
            scratch_cap = k;
            scratch = (const phjo_tuple**)realloc((void*)scratch, scratch_cap * sizeof(*scratch));
            k = phjo_table_get_all(table, S[i].id, scratch, scratch_cap);
        }
        for (size_t j = 0; j != k; ++j, ++n)
            if (out && n < cap) {
                out[n].id = S[i].id;
                out[n].payloadA = scratch[j]->payload;
                out[n].payloadB = S[i].payload;
            }
    }
    free((void*)scratch);
    phjo_table_free(table);
    return n;
}

/* HashJoiner::GetPartitioningConfiguration (src/RadixCluster/HashJoin.hpp:149-188), including its
 * quirk: when a batch falls below MinBatchSize the worker count is re-derived from that relation
 * and then applies to BOTH relations, while the other relation keeps its old batch size. */
void phjo_partitioning_configuration(size_t sizeA, size_t sizeB, size_t pool_workers,
                                     size_t min_batch, size_t* workers, size_t* batchA,
                                     size_t* batchB) {
    size_t w = pool_workers;
    size_t a = (size_t)((double)sizeA / (double)w);
    size_t b = (size_t)((double)sizeB / (double)w);
    if (a < min_batch) {
        w = (size_t)ceil((double)sizeA / (double)min_batch);
        a = min_batch;
    }
    if (b < min_batch) {
        w = (size_t)ceil((double)sizeB / (double)min_batch);
        b = min_batch;
    }
    *workers = w;
    *batchA = a;
    *batchB = b;
}

/* HashJoiner::Partition (src/RadixCluster/HashJoin.hpp:333-440): pipeline of
 *   scanTable (:343-357)            per-worker histogram of Hash(id, P)
 *   createPrefixSumTable (:363-390) per-partition exclusive scan over workers (+ partition sizes)
 *   ComputePartitionsBoundaries (:18-25) running (first, second) pairs
 *   partitionTable (:394-412)       out[bounds(p).first + cursor(p, w)++] = in[i]
 * with worker w owning [batch*w, batch*(w+1)) and the last worker running to n (:416-426). */
void phjo_radix_partition(const phjo_tuple* in, size_t n, size_t P, int hash_id, uint64_t seed,
                          size_t workers, size_t batch, phjo_tuple* out, uint64_t* bounds,
                          uint64_t* prefix_out) {
    /* PrefixSumTable: worker-major, m_table[worker * P + partition] (:46-47) */
    uint64_t* table = (uint64_t*)calloc((workers != 0 && P != 0) ? workers * P : 1, sizeof(uint64_t));
    for (size_t w = 0; w != workers; ++w) {
        size_t start = batch * w, end = (w + 1 == workers) ? n : batch * (w + 1);
        if (start > n) start = n; /* the reference would index out of range; never hit by callers */
        if (end > n) end = n;
        for (size_t i = start; i < end; ++i) table[w * P + phjo_hash(hash_id, in[i].id, seed, P)]++;
    }
    uint64_t running_end = 0;
    for (size_t p = 0; p != P; ++p) {
        uint64_t running = 0;
        for (size_t w = 0; w != workers; ++w) { /* :367-381 */
            uint64_t current = table[w * P + p];
            table[w * P + p] = running;
            running += current;
        }
        bounds[2 * p] = running_end; /* :18-25 */
        bounds[2 * p + 1] = running_end + running;
        running_end += running;
    }
    if (prefix_out) memcpy(prefix_out, table, workers * P * sizeof(uint64_t));
    for (size_t w = 0; w != workers; ++w) { /* :400-407 */
        size_t start = batch * w, end = (w + 1 == workers) ? n : batch * (w + 1);
        if (start > n) start = n;
        if (end > n) end = n;
        for (size_t i = start; i < end; ++i) {
            uint64_t p = phjo_hash(hash_id, in[i].id, seed, P);
            out[bounds[2 * p] + table[w * P + p]++] = in[i];
        }
    }
    free(table);
}

/* HashJoiner::Run + Join (src/RadixCluster/HashJoin.hpp:190-241, 243-331). */
int phjo_join_radix(const phjo_tuple* R, size_t nR, const phjo_tuple* S, size_t nS, size_t P,
                    size_t pool_workers, int table_kind, int hash_id, uint64_t seed_partition,
                    uint64_t seed_table, phjo_result* out) {
    memset(out, 0, sizeof(*out));
    if (P == 0 || pool_workers == 0) return 1;
    size_t workers, batchA, batchB;
    phjo_partitioning_configuration(nR, nS, pool_workers, 10000, &workers, &batchA, &batchB);
    if (workers == 0) workers = 1; /* both relations empty: nothing to do */

    phjo_tuple* pR = (phjo_tuple*)malloc((nR ? nR : 1) * sizeof(phjo_tuple)); /* :195-198 */
    phjo_tuple* pS = (phjo_tuple*)malloc((nS ? nS : 1) * sizeof(phjo_tuple));
    uint64_t* bR = (uint64_t*)malloc(2 * P * sizeof(uint64_t));
    uint64_t* bS = (uint64_t*)malloc(2 * P * sizeof(uint64_t));

    uint64_t t0 = now_ns(); /* :208 */
    phjo_radix_partition(R, nR, P, hash_id, seed_partition, workers, batchA, pR, bR, NULL);
    phjo_radix_partition(S, nS, P, hash_id, seed_partition, workers, batchB, pS, bS, NULL);
    uint64_t t1 = now_ns(); /* :224 */
    out->partition_ns = t1 - t0;

    int rc = 0;
    uint64_t joined = 0;
    /* worker id handles partitions id, id + W, ... (:268-269); order is irrelevant to the sum */
    for (size_t p = 0; p != P && rc == 0; ++p) {
        uint64_t a0 = bR[2 * p], a1 = bR[2 * p + 1];
        if (a1 - a0 == 0) continue; /* :273-276 */
        uint64_t tb = now_ns();
        phjo_table* table = phjo_table_new(table_kind, 0, hash_id, seed_table, a1 - a0); /* :278 */
        for (uint64_t i = a0; i != a1; ++i) /* :283-286 */
            if (phjo_table_insert(table, pR[i].id, &pR[i])) rc = 2;
        uint64_t tp = now_ns();
        for (uint64_t i = bS[2 * p]; i != bS[2 * p + 1]; ++i) /* :295-301 */
            if (phjo_table_get(table, pS[i].id) != NULL) joined++;
        out->build_ns += tp - tb;
        out->probe_ns += now_ns() - tp;
        phjo_table_free(table);
    }
    out->matches = joined;
    free(pR);
    free(pS);
    free(bR);
    free(bS);
    return rc;
}

/* Independent cross-check: sort R's keys, binary-search every S key. */
static int cmp_i64(const void* a, const void* b) {
    int64_t x = *(const int64_t*)a, y = *(const int64_t*)b;
    return (x > y) - (x < y);
}
uint64_t phjo_count_by_sort(const phjo_tuple* R, size_t nR, const phjo_tuple* S, size_t nS) {
    if (nR == 0) return 0;
    int64_t* keys = (int64_t*)malloc(nR * sizeof(int64_t));
    for (size_t i = 0; i != nR; ++i) keys[i] = R[i].id;
    qsort(keys, nR, sizeof(int64_t), cmp_i64);
    uint64_t count = 0;
    for (size_t i = 0; i != nS; ++i) {
        int64_t k = S[i].id;
        size_t lo = 0, hi = nR;
        while (lo < hi) {
            size_t mid = lo + (hi - lo) / 2;
            if (keys[mid] < k) {
                lo = mid + 1;
            } else {
                hi = mid;
            }
        }
        if (lo < nR && keys[lo] == k) count++;
    }
    free(keys);
    return count;
}
