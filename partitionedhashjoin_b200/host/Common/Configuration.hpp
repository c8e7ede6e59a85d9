// CLI-facing configuration: the enum spellings, defaults and validation rules of the reference
// (src/Common/Configuration.hpp:12-68, src/Common/Configuration.cpp:4-106, src/Arguments.hpp:7-19),
// plus the options that only exist for the GPU engine (hash function, seeds, devices).
#pragma once
#include <cstddef>
#include <cstdint>
#include <stdexcept>
#include <string>

namespace NoPartitioning {
struct Configuration {
    size_t MinBatchSize = 10000;  // CPU batching knob of the reference; accepted, unused on the GPU
};
}  // namespace NoPartitioning

namespace RadixClustering {
struct Configuration {
    size_t MinBatchSize = 10000;
    size_t NumberOfPartitions = 32;
};
}  // namespace RadixClustering

namespace Common {

// "RadixParitioning" (sic) is the reference's spelling in the enum and in its JSON output; it is
// kept so that consumers of that output keep working.
enum class JoinAlgorithmType : uint8_t { NoPartitioning = 0, RadixParitioning = 1 };
enum class ResultsFormat : uint8_t { JSON = 0 };
enum class OutputType : uint8_t { File = 0 };
enum SeverityLevel { trace, debug, info, error, critical };
enum class HashFunction : uint8_t { XXH3 = 0, Murmur3 = 1, City = 2 };

inline JoinAlgorithmType GetJoinAlgorithmTypeFromString(const std::string& s) {
    if (s == "no-partitioning") return JoinAlgorithmType::NoPartitioning;
    if (s == "radix-partitioning") return JoinAlgorithmType::RadixParitioning;
    throw std::runtime_error("Unrecognized join algorithm type: " + s + ".");
}
inline const char* ToString(JoinAlgorithmType t) {
    return t == JoinAlgorithmType::NoPartitioning ? "no-partitioning" : "radix-partitioning";
}
inline ResultsFormat GetResultsFormatFromString(const std::string& s) {
    if (s == "json") return ResultsFormat::JSON;
    throw std::runtime_error("Unrecognized results format: " + s + ".");
}
inline OutputType GetOutputTypeFromString(const std::string& s) {
    if (s == "file") return OutputType::File;
    throw std::runtime_error("Unrecognized output type: " + s + ".");
}
inline SeverityLevel SeverityLevelFromString(const std::string& s) {
    static const char* names[] = {"trace", "debug", "info", "error", "critical"};
    for (int i = 0; i < 5; ++i)
        if (s == names[i]) return static_cast<SeverityLevel>(i);
    throw std::runtime_error("Unrecognized logger level: " + s + ".");
}
inline HashFunction GetHashFunctionFromString(const std::string& s) {
    if (s == "xxh3" || s == "xxhash") return HashFunction::XXH3;
    if (s == "murmur3") return HashFunction::Murmur3;
    if (s == "city") return HashFunction::City;
    throw std::runtime_error("Unrecognized hash function: " + s + ".");
}
inline const char* ToString(HashFunction h) {
    return h == HashFunction::XXH3 ? "xxh3" : h == HashFunction::Murmur3 ? "murmur3" : "city";
}

struct FileConfiguration {
    std::string Name = "hashjoin.txt";
};

struct OutputConfiguration {
    OutputType Type = OutputType::File;
    FileConfiguration File;
    void Validate() const {
        if (Type == OutputType::File && File.Name.empty())
            throw std::invalid_argument("OutputConfiguration::Validate: empty configuration filename specified.");
    }
};

struct ResultsFormatConfiguration {
    ResultsFormat Format = ResultsFormat::JSON;
    std::string TimeUnit = "ms";
    void Validate() const {
        for (const char* u : {"ns", "us", "ms", "s"})
            if (TimeUnit == u) return;
        throw std::invalid_argument("ResultsFormatConfiguration::Validate: Unrecognized time unit: " + TimeUnit);
    }
};

struct LoggerConfiguration {
    SeverityLevel LogLevel = debug;  // the CLI default (reference src/main.cpp:158-161)
};

// Options with no counterpart in the reference.
struct GpuConfiguration {
    HashFunction Hash = HashFunction::XXH3;
    uint64_t Seed = 0x9E3779B97F4A7C15ULL;  // partition-hasher seed; the reference draws a random one
    uint64_t TableSeed = 1;
    int64_t DataSeed = 12345;   // base seed of the Zipf generator's batches
    size_t DataBatches = 64;    // generator batches (fixed so the data does not depend on the host)
    int Device = 0;
    int Gpus = 1;
    unsigned RadixBits[2] = {0, 0};
    int Repeat = 1;             // joins per run; the best is reported
    bool NoTmaStore = false;
    bool ChainedTable = false;  // no-partitioning join: bucket-chained table (SeparateChaining) instead of
                                // the open-addressing one (LinearProbing)
    bool Materialize = false;   // fill the returned Table<JoinedTuple> (the reference leaves it empty)
    bool StreamUpload = false;  // Run() through phj_join_host: the upload of tableB overlaps the join of the
                                // chunks that have landed (phase timings become sums over the chunks)
};

struct Configuration {
    JoinAlgorithmType JoinType = JoinAlgorithmType::NoPartitioning;
    ResultsFormatConfiguration OutputFormatConfig;
    OutputConfiguration OutputConfig;
    size_t PrimaryRelationSize = 10'000'000;
    size_t SecondaryRelationSize = 200'000'000;
    double SkewParameter = 1.05;
    NoPartitioning::Configuration NoPartitioningConfig;
    RadixClustering::Configuration RadixClusteringConfig;
    LoggerConfiguration LoggerConfig;
    GpuConfiguration Gpu;
};

}  // namespace Common
