// oracle/ref_harness.cpp -- TEST INFRASTRUCTURE ONLY (never linked into the product).
//
// A C-ABI wrapper around the UNMODIFIED reference sources under /root/reference/src, compiled
// where they lie by oracle/Makefile into oracle/_ref/libphj_ref.so. It instantiates the joiners
// exactly as the reference's own main does (src/main.cpp:211-217 for the hash-table factory,
// :91-92 for NoPartitioning::HashJoiner, :120-121 for RadixClustering::HashJoiner) and exposes
//   * the reference DataGenerator (Sequential / Zipf) and LCG with an explicit seed,
//   * XXH3 as called by src/Common/XXHasher.hpp:19-22,
//   * both joiners, returning the match count the reference only logs,
//   * the reference hash tables' Insert / Exists / Get / GetAll,
// so that tests/ can pin oracle/phj_oracle.c (the C restatement) and the CUDA path against the
// real thing, and bench.py can time the reference CPU joins as the `cpu_baseline`.
//
// Nothing in here is copied from the reference: it only *calls* it.

#include <atomic>
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <mutex>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>

#include <sys/wait.h>
#include <unistd.h>

#include "Common/Configuration.hpp"
#include "Common/IHasher.hpp"
#include "Common/IThreadPool.hpp"
#include "Common/Logger.hpp"
#include "Common/Random.hpp"
#include "Common/Results.hpp"
#include "Common/Table.hpp"
#include "Common/ThreadPool.hpp"
#include "Common/XXHasher.hpp"
#include "DataGenerator/Sequential.hpp"
#include "DataGenerator/Zipf.hpp"
#include "HashTables/LinearProbing.hpp"
#include "HashTables/SeparateChaining.hpp"
#include "NoPartitioning/HashJoin.hpp"
#include "RadixCluster/HashJoin.hpp"

// ------------------------------------------------------------------------------------------------
// Replacements for the two reference translation units that need Boost (Logger.cpp, Table.cpp).
// ------------------------------------------------------------------------------------------------
namespace {
std::atomic<uint64_t> g_joined_max{0};
std::atomic<int> g_joined_seen{0};
std::atomic<int> g_echo_level{99};  // lines with severity >= this go to stderr
}  // namespace

namespace phj_shim {
void sink(int severity, const std::string& line) {
    // NPJ:   "Joined <n> tuples."   (src/NoPartitioning/HashJoin.hpp:184)
    // radix: "Joined  <n> tuples"   (src/RadixCluster/HashJoin.hpp:320-321) -- may be emitted by
    // more than one worker and an early emission can miss a straggler's fetch_add (the check at
    // :316 races with :311), hence the running maximum.
    if (line.compare(0, 7, "Joined ") == 0) {
        const char* p = line.c_str() + 7;
        while (*p == ' ') ++p;
        uint64_t n = std::strtoull(p, nullptr, 10);
        uint64_t prev = g_joined_max.load();
        while (n > prev && !g_joined_max.compare_exchange_weak(prev, n)) {
        }
        g_joined_seen.fetch_add(1);
    }
    if (severity >= g_echo_level.load()) {
        std::fprintf(stderr, "[phj_ref %d] %s\n", severity, line.c_str());
    }
}
}  // namespace phj_shim

namespace Common {
SeverityLevel SeverityLevelFromString(const std::string& level) {
    static const char* names[] = {"trace", "debug", "info", "error", "critical"};
    for (int i = 0; i != 5; ++i) {
        if (level == names[i]) return static_cast<SeverityLevel>(i);
    }
    throw std::runtime_error("Unrecognized logger level: " + level + ".");
}
std::istream& operator>>(std::istream& in, SeverityLevel& obj) {
    std::string s;
    in >> s;
    obj = SeverityLevelFromString(s);
    return in;
}
void InitializeLogger(const LoggerConfiguration&) {}
LoggerType GetNewLogger() { return LoggerType{}; }
void AddComponentAttributeToLogger(LoggerType&, std::string) {}
void AddTableIDToLogger(LoggerType&, std::string) {}
std::string generate_uuid() {
    static std::atomic<uint64_t> counter{0};
    return "phj-ref-" + std::to_string(counter.fetch_add(1));
}
std::ostream& operator<<(std::ostream& out, const Tuple& tuple) {
    return out << tuple.id << ", " << tuple.payload;
}
std::ostream& operator<<(std::ostream& out, const JoinedTuple& tuple) {
    return out << tuple.id << ", " << tuple.payloadA << ", " << tuple.payloadB;
}
}  // namespace Common

// ------------------------------------------------------------------------------------------------
// Harness-side helpers
// ------------------------------------------------------------------------------------------------
namespace {
using TableT = Common::Table<Common::Tuple>;

// Same arithmetic as Common::XXHasher::Hash (src/Common/XXHasher.hpp:19-22) but with a settable
// seed: XXHasher::m_seed is private, random and has no setter (:12-17,27). The joiners and tables
// are duck-typed on Hash(key, cardinality), so this is a legal HasherType.
class SeededXXHasher : public Common::IHasher {
   public:
    explicit SeededXXHasher(uint64_t seed) : m_seed(seed) {}
    uint64_t Hash(int64_t key, size_t cardinality) override {
        auto hash = XXH3_64bits_withSeed(static_cast<const void*>(&key), sizeof(key), m_seed);
        return hash % cardinality;
    }

   private:
    uint64_t m_seed;
};

// Reaches the protected Zipf::generate the same way tests/DataGenerator/ZipfTest.hpp:7-13 does.
class ZipfAccess : public DataGenerator::Zipf {
   public:
    static uint64_t Generate(double alpha, uint64_t cardinality,
                             std::shared_ptr<Common::IRandomNumberGenerator> generator) {
        return DataGenerator::Zipf::generate(alpha, cardinality, generator);
    }
};

std::shared_ptr<Common::IThreadPool> make_pool(int threads) {
    if (threads <= 0) {
        // src/main.cpp:235-241
        threads = static_cast<int>(std::thread::hardware_concurrency()) - 1;
        if (threads < 1) threads = 1;
    }
    return std::make_shared<Common::ThreadPool>(static_cast<size_t>(threads));
}

thread_local std::string g_error;

struct RefTable {
    std::shared_ptr<TableT> table;
};
}  // namespace

extern "C" {

struct phjref_tuple {
    int64_t id;
    int64_t payload;
};

struct phjref_result {
    uint64_t matches;       // max over captured "Joined" records
    uint64_t partition_ns;  // as reported through the reference's own HashJoinTimer
    uint64_t build_ns;
    uint64_t probe_ns;      // NPJ: includes build (src/Common/Results.hpp:202)
    uint64_t wall_ns;       // around Run(), i.e. including the reference's internal allocations
    uint32_t workers;
    uint32_t joined_records;
};

const char* phjref_last_error(void) { return g_error.c_str(); }

int phjref_default_workers(void) {
    int t = static_cast<int>(std::thread::hardware_concurrency()) - 1;
    return t < 1 ? 1 : t;
}

void phjref_set_echo_level(int level) { g_echo_level.store(level); }

uint64_t phjref_xxh3_64(int64_t key, uint64_t seed) {
    return XXH3_64bits_withSeed(static_cast<const void*>(&key), sizeof(key), seed);
}

// Common::XXHasher-equivalent Hash(key, cardinality) with an explicit seed.
uint64_t phjref_hash(int64_t key, uint64_t seed, uint64_t cardinality) {
    SeededXXHasher h(seed);
    return h.Hash(key, cardinality);
}

void phjref_lcg_doubles(long seed, size_t n, double* out) {
    Common::MultiplicativeLCGRandomNumberGeneratorFactory factory;
    auto gen = factory.GetNewGenerator(seed);
    for (size_t i = 0; i != n; ++i) out[i] = gen->Next();
}

int phjref_zipf_samples(double alpha, uint64_t cardinality, long seed, size_t n, uint64_t* out) {
    try {
        Common::MultiplicativeLCGRandomNumberGeneratorFactory factory;
        auto gen = factory.GetNewGenerator(seed);
        for (size_t i = 0; i != n; ++i) out[i] = ZipfAccess::Generate(alpha, cardinality, gen);
        return 0;
    } catch (std::exception& e) {
        g_error = e.what();
        return 1;
    }
}

// ---- tables -------------------------------------------------------------------------------------
void* phjref_table_new(size_t n) {
    auto* t = new RefTable;
    t->table = std::make_shared<TableT>(n, Common::generate_uuid());
    return t;
}
void phjref_table_free(void* h) { delete static_cast<RefTable*>(h); }
size_t phjref_table_size(void* h) { return static_cast<RefTable*>(h)->table->GetSize(); }
phjref_tuple* phjref_table_data(void* h) {
    auto& t = *static_cast<RefTable*>(h)->table;
    return t.GetSize() ? reinterpret_cast<phjref_tuple*>(&t[0]) : nullptr;
}

// Sequential::FillTable exactly as generateTables does (src/main.cpp:54-56).
int phjref_fill_sequential(void* h, int64_t start, int threads) {
    try {
        auto pool = make_pool(threads);
        auto fut = DataGenerator::Sequential::FillTable(
            pool, static_cast<RefTable*>(h)->table, DataGenerator::Sequential::Parameters{start});
        fut.wait();
        auto errs = fut.get();
        pool->Stop();
        if (!errs.Empty()) {
            g_error = "Sequential::FillTable task failed";
            return 1;
        }
        return 0;
    } catch (std::exception& e) {
        g_error = e.what();
        return 1;
    }
}

// Seeded Zipf fill. Zipf::FillTable seeds every batch from std::random_device and its batch
// boundaries depend on the worker count (src/DataGenerator/Zipf.cpp:70-78,86), so a reproducible
// table needs the split done here: `batches` equal slices (the last takes the remainder, like
// :99-101), slice b drawing from GetNewGenerator(base_seed + b) (src/Common/Random.cpp:38-41) and
// each element produced by the reference's own Zipf::generate + the :89-92 id/payload rule.
int phjref_fill_zipf_seeded(void* h, double alpha, int64_t range_first, int64_t range_second,
                            long base_seed, size_t batches, int threads) {
    try {
        auto table = static_cast<RefTable*>(h)->table;
        if (range_first >= range_second) throw std::invalid_argument("bad Zipf range");
        if (batches == 0) throw std::invalid_argument("batches must be > 0");
        const size_t size = table->GetSize();
        const size_t batchSize = size / batches;
        const int64_t cardinality = range_second - range_first + 1;
        const int64_t correction = range_first - 1;
        auto factory = std::make_shared<Common::MultiplicativeLCGRandomNumberGeneratorFactory>();
        auto pool = make_pool(threads);
        std::vector<std::function<void()>> tasks;
        for (size_t b = 0; b != batches; ++b) {
            size_t start = batchSize * b;
            size_t end = (b + 1 == batches) ? size : batchSize * (b + 1);
            long seed = base_seed + static_cast<long>(b);
            tasks.push_back([=]() {
                auto gen = factory->GetNewGenerator(seed);
                for (size_t i = start; i != end; ++i) {
                    (*table)[i].id = ZipfAccess::Generate(alpha, cardinality, gen) + correction;
                    (*table)[i].payload = i;
                }
            });
        }
        auto fut = pool->Push(std::move(tasks));
        fut.wait();
        auto errs = fut.get();
        pool->Stop();
        if (!errs.Empty()) {
            g_error = "Zipf fill task failed (alpha < 0.01?)";
            return 1;
        }
        return 0;
    } catch (std::exception& e) {
        g_error = e.what();
        return 1;
    }
}

// ---- joins --------------------------------------------------------------------------------------
// algo: 0 = NoPartitioning::HashJoiner, 1 = RadixClustering::HashJoiner.
// table_kind: 0 = LinearProbingFactory<Tuple,3,H> (what main wires, src/main.cpp:216-217),
//             1 = SeparateChainingFactory<Tuple,3,H> (only exercised by the reference's tests).
// seeded != 0 uses SeededXXHasher(seed_partition / seed_table); seeded == 0 uses the reference's
// own Common::XXHasher (random seeds), i.e. the completely stock configuration.
}  // extern "C"

namespace {
template <typename Hasher, typename Factory>
int run_join(std::shared_ptr<TableT> R, std::shared_ptr<TableT> S, int algo, size_t partitions,
             int threads, const Hasher& partitionHasher, const Factory& factory,
             phjref_result* out) {
    auto pool = make_pool(threads);
    Common::Parameters params;
    std::shared_ptr<Common::IHashJoinTimer> timer = std::make_shared<Common::HashJoinTimer>(params);
    g_joined_max.store(0);
    g_joined_seen.store(0);
    auto t0 = std::chrono::steady_clock::now();
    if (algo == 0) {
        NoPartitioning::Configuration cfg{};
        auto joiner = NoPartitioning::HashJoiner<Factory>(cfg, pool, factory);
        joiner.Run(R, S, timer);
    } else {
        RadixClustering::Configuration cfg{};
        cfg.NumberOfPartitions = partitions;
        auto joiner = RadixClustering::HashJoiner<Factory, Hasher>(cfg, pool, partitionHasher,
                                                                    factory);
        joiner.Run(R, S, timer);
    }
    auto t1 = std::chrono::steady_clock::now();
    out->workers = static_cast<uint32_t>(pool->GetNumberOfWorkers());
    pool->Stop();  // joins every worker, so all "Joined" records have been emitted
    auto res = timer->GetResult();
    out->matches = g_joined_max.load();
    out->joined_records = static_cast<uint32_t>(g_joined_seen.load());
    out->partition_ns = static_cast<uint64_t>(res.GetPartitioningPhaseDuration().count());
    out->build_ns = static_cast<uint64_t>(res.GetBuildPhaseDuration().count());
    out->probe_ns = static_cast<uint64_t>(res.GetProbePhaseDuration().count());
    out->wall_ns = static_cast<uint64_t>(
        std::chrono::duration_cast<std::chrono::nanoseconds>(t1 - t0).count());
    return 0;
}

template <typename Hasher>
int run_join_tables(std::shared_ptr<TableT> R, std::shared_ptr<TableT> S, int algo, int table_kind,
                    size_t partitions, int threads, const Hasher& partitionHasher,
                    const Hasher& tableHasher, phjref_result* out) {
    if (table_kind == 0) {
        HashTables::LinearProbingFactory<Common::Tuple, 3, Hasher> factory(
            HashTables::LinearProbingConfiguration{}, tableHasher);
        return run_join(R, S, algo, partitions, threads, partitionHasher, factory, out);
    }
    HashTables::SeparateChainingFactory<Common::Tuple, 3, Hasher> factory(
        HashTables::SeparateChainingConfiguration{}, tableHasher);
    return run_join(R, S, algo, partitions, threads, partitionHasher, factory, out);
}
}  // namespace

extern "C" {

int phjref_join_tables(void* hR, void* hS, int algo, int table_kind, size_t partitions, int threads,
                       int seeded, uint64_t seed_partition, uint64_t seed_table,
                       phjref_result* out) {
    std::memset(out, 0, sizeof(*out));
    try {
        auto R = static_cast<RefTable*>(hR)->table;
        auto S = static_cast<RefTable*>(hS)->table;
        if (seeded) {
            return run_join_tables(R, S, algo, table_kind, partitions, threads,
                                   SeededXXHasher(seed_partition), SeededXXHasher(seed_table), out);
        }
        return run_join_tables(R, S, algo, table_kind, partitions, threads, Common::XXHasher{},
                               Common::XXHasher{}, out);
    } catch (std::exception& e) {
        g_error = std::string("reference join threw: ") + e.what();
        return 1;
    }
}

// The same join in a forked child. RadixClustering::HashJoiner::Run never releases its two
// partitioned tables (3.36 GB per 10 M x 200 M join stay resident after Run returns -- measured
// here: RSS grows by exactly that per call), so a benchmark loop of K joins in one process needs
// K x 3.36 GB. The child shares the input tables copy-on-write, runs the UNMODIFIED joiner with the
// reference's own phase timers, sends the result through a pipe and exits; the parent's memory does
// not grow. Timing is the reference's own (IHashJoinTimer), so the fork does not enter it.
int phjref_join_tables_forked(void* hR, void* hS, int algo, int table_kind, size_t partitions,
                              int threads, int seeded, uint64_t seed_partition, uint64_t seed_table,
                              phjref_result* out) {
    std::memset(out, 0, sizeof(*out));
    int fd[2];
    if (pipe(fd) != 0) {
        g_error = "pipe() failed";
        return 1;
    }
    std::fflush(nullptr);
    const pid_t pid = fork();
    if (pid < 0) {
        close(fd[0]);
        close(fd[1]);
        g_error = "fork() failed";
        return 1;
    }
    if (pid == 0) {
        close(fd[0]);
        phjref_result r;
        int rc = phjref_join_tables(hR, hS, algo, table_kind, partitions, threads, seeded,
                                    seed_partition, seed_table, &r);
        ssize_t w = write(fd[1], &rc, sizeof(rc));
        w += write(fd[1], &r, sizeof(r));
        (void)w;
        _exit(0);
    }
    close(fd[1]);
    int rc = 1;
    phjref_result r;
    std::memset(&r, 0, sizeof(r));
    bool ok = read(fd[0], &rc, sizeof(rc)) == (ssize_t)sizeof(rc) &&
              read(fd[0], &r, sizeof(r)) == (ssize_t)sizeof(r);
    close(fd[0]);
    int status = 0;
    waitpid(pid, &status, 0);
    if (!ok) {
        g_error = "the forked reference join died before reporting (out of memory?)";
        return 1;
    }
    *out = r;
    if (rc) g_error = "reference join failed in the forked child";
    return rc;
}

// Convenience: copy caller arrays into reference tables, then join.
int phjref_join(const phjref_tuple* R, size_t nR, const phjref_tuple* S, size_t nS, int algo,
                int table_kind, size_t partitions, int threads, int seeded, uint64_t seed_partition,
                uint64_t seed_table, phjref_result* out) {
    void* hR = phjref_table_new(nR);
    void* hS = phjref_table_new(nS);
    if (nR) std::memcpy(phjref_table_data(hR), R, nR * sizeof(phjref_tuple));
    if (nS) std::memcpy(phjref_table_data(hS), S, nS * sizeof(phjref_tuple));
    int rc = phjref_join_tables(hR, hS, algo, table_kind, partitions, threads, seeded,
                                seed_partition, seed_table, out);
    phjref_table_free(hR);
    phjref_table_free(hS);
    return rc;
}

// ---- hash tables (the reference's own unit-test surface, tests/NoPartitioningHashJoin/
// HashTableTest.hpp:10-82) -------------------------------------------------------------------------
// Inserts build[0..nb) (value = &build[i]) from `insert_threads` threads, then for every probe key
// reports Exists, the index of the tuple Get returned (-1 = nullptr) and GetAll().size().
int phjref_table_probe(int table_kind, double size_ratio, uint64_t seed, size_t number_of_objects,
                       const phjref_tuple* build, size_t nb, int insert_threads,
                       const int64_t* probe, size_t np, uint8_t* exists, int64_t* get_index,
                       uint64_t* getall_count) {
    try {
        const Common::Tuple* base = reinterpret_cast<const Common::Tuple*>(build);
        auto run = [&](auto table) {
            auto inserter = [&](size_t a, size_t b) {
                for (size_t i = a; i != b; ++i) table->Insert(base[i].id, &base[i]);
            };
            if (insert_threads <= 1) {
                inserter(0, nb);
            } else {
                std::vector<std::thread> ts;
                size_t per = nb / insert_threads;
                for (int t = 0; t != insert_threads; ++t) {
                    size_t a = per * t, b = (t + 1 == insert_threads) ? nb : per * (t + 1);
                    ts.emplace_back(inserter, a, b);
                }
                for (auto& t : ts) t.join();
            }
            for (size_t i = 0; i != np; ++i) {
                exists[i] = table->Exists(probe[i]) ? 1 : 0;
                const Common::Tuple* g = table->Get(probe[i]);
                get_index[i] = g ? static_cast<int64_t>(g - base) : -1;
                getall_count[i] = table->GetAll(probe[i]).size();
            }
        };
        SeededXXHasher hasher(seed);
        if (table_kind == 0) {
            HashTables::LinearProbingConfiguration cfg{};
            if (size_ratio > 0) cfg.HASH_TABLE_SIZE_RATIO = size_ratio;
            HashTables::LinearProbingFactory<Common::Tuple, 3, SeededXXHasher> f(cfg, hasher);
            run(f.New(number_of_objects));
        } else {
            HashTables::SeparateChainingConfiguration cfg{};
            if (size_ratio > 0) cfg.HASH_TABLE_SIZE_RATIO = size_ratio;
            HashTables::SeparateChainingFactory<Common::Tuple, 3, SeededXXHasher> f(cfg, hasher);
            run(f.New(number_of_objects));
        }
        return 0;
    } catch (std::exception& e) {
        g_error = e.what();
        return 1;
    }
}

// Bucket counts chosen by the reference for n objects (src/HashTables/LinearProbing.cpp:7-12,
// src/HashTables/SeparateChaining.cpp:7-13).
uint64_t phjref_number_of_buckets(int table_kind, double size_ratio, uint64_t n) {
    if (table_kind == 0) {
        HashTables::LinearProbingConfiguration cfg{};
        if (size_ratio > 0) cfg.HASH_TABLE_SIZE_RATIO = size_ratio;
        return HashTables::internal::LinearProbing::getNumberOfBuckets(cfg, n);
    }
    HashTables::SeparateChainingConfiguration cfg{};
    if (size_ratio > 0) cfg.HASH_TABLE_SIZE_RATIO = size_ratio;
    return HashTables::internal::SeparateChaining::getNumberOfBuckets(cfg, n);
}

// The reference's JSON rendering of a timing result (src/Common/Results.hpp:262-279) for given
// parameters, so the product CLI's output can be diffed against it.
int phjref_format_json(const char* const* keys, const char* const* values, size_t nparams,
                       uint64_t partition_ns, uint64_t build_ns, uint64_t probe_ns,
                       const char* unit, char* out, size_t out_cap) {
    try {
        Common::Parameters params;
        for (size_t i = 0; i != nparams; ++i) params.SetParameter(keys[i], values[i]);
        Common::HashJoinTimingResult r(std::chrono::nanoseconds(build_ns),
                                       std::chrono::nanoseconds(probe_ns),
                                       std::chrono::nanoseconds(partition_ns), params);
        Common::ResultsFormatConfiguration cfg{Common::ResultsFormat::JSON, unit};
        Common::JSONResultsFormatter fmt(cfg);
        std::ostringstream os;
        fmt.Format(os, r);
        std::string s = os.str();
        if (s.size() + 1 > out_cap) {
            g_error = "output buffer too small";
            return 1;
        }
        std::memcpy(out, s.c_str(), s.size() + 1);
        return 0;
    } catch (std::exception& e) {
        g_error = e.what();
        return 1;
    }
}

// String <-> enum spellings of the CLI (src/Common/Configuration.cpp:4-22).
int phjref_join_type_from_string(const char* s) {
    try {
        return static_cast<int>(Common::GetJoinAlgorithmTypeFromString(s));
    } catch (std::exception& e) {
        g_error = e.what();
        return -1;
    }
}

}  // extern "C"
