// Command-line surface of phjoin: the reference's flags, spellings, defaults and validation
// (reference src/main.cpp:141-208, src/Arguments.hpp:7-19) parsed without Boost, plus the flags
// that only make sense for the GPU engine (--hash, --seed, --table-seed, --data-seed,
// --radix-bits, --device, --gpus, --repeat, --no-tma-store, --materialize, --stream-upload).
#pragma once
#include <cstdlib>
#include <iostream>
#include <map>
#include <set>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

#include "Common/Configuration.hpp"

namespace Arguments {

inline std::string HelpText() {
    std::ostringstream o;
    o << "Allowed options:\n"
         "  -h [ --help ]                         Help screen\n"
         "  --primary arg (=10000000)             Size of the primary relation.\n"
         "  --secondary arg (=200000000)          Size of the secondary relation.\n"
         "  --skew arg (=1.05)                    Parameter of skew for Zipf distribution used for the\n"
         "                                        generation of tuples for secondary relation.\n"
         "  --log arg (=debug)                    Logging level. One of {trace, debug, info, error, critical}.\n"
         "  --join arg                            Type of join algorithm: either no-partitioning or\n"
         "                                        radix-partitioning.\n"
         "  --format arg (=json)                  Format of the output. Currently only JSON is supported.\n"
         "  -u [ --unit ] arg (=ms)               Duration unit of the timing output. One of {ns, us, ms, s}.\n"
         "  -o [ --output ] arg (=file)           Type of the output. Currently only file is supported.\n"
         "  -f [ --filename ] arg (=hashjoin.txt) Name of the file if output type is file.\n"
         "  -p [ --partitions ] arg               Number of partitions for algorithms using partitioning.\n"
         "GPU engine options (not in the reference):\n"
         "  --hash arg (=xxh3)                    Partitioning hash: xxh3 (XXH3_64bits_withSeed, the reference's),\n"
         "                                        murmur3 (MurmurHash3_x64_128 low word), city (CityHash64WithSeed,\n"
         "                                        v1.0 short-key path).\n"
         "  --seed arg                            Seed of the partitioning hasher.\n"
         "  --table-seed arg                      Seed of the hash-table hasher.\n"
         "  --data-seed arg (=12345)              Base seed of the Zipf generator (batch b uses seed + b).\n"
         "  --radix-bits arg                      b1,b2: bits of partitioning pass 1 and 2.\n"
         "  --device arg (=0)                     CUDA device.\n"
         "  --gpus arg (=1)                       Number of GPUs (1 .. 16). Radix join: sharded by partition,\n"
         "                                        --partitions then = GPUs x partitions per GPU (<= 256;\n"
         "                                        default: the largest multiple of the GPUs up to 64).\n"
         "                                        No-partitioning join: every GPU builds the whole table and\n"
         "                                        probes its share of the secondary relation.\n"
         "  --repeat arg (=1)                     Joins per run; the fastest is reported.\n"
         "  --no-tma-store                        Flush scatter tiles with st.global instead of TMA bulk stores.\n"
         "  --table arg (=linear-probing)         Hash table of the no-partitioning join: linear-probing or\n"
         "                                        separate-chaining (the reference's two HashTables).\n"
         "  --materialize                         Radix join only: also produce the joined table {id, payloadA,\n"
         "                                        payloadB} (the reference returns it empty) and log its size.\n"
         "  --stream-upload                       Overlap the host-to-device upload with the join: the secondary\n"
         "                                        relation goes up in chunks that are joined as they land; phase\n"
         "                                        timings are then sums over the chunks and the end-to-end time\n"
         "                                        is logged.\n";
    return o.str();
}

inline uint64_t ParseUnsigned(const std::string& name, const std::string& v) {
    size_t pos = 0;
    unsigned long long x = 0;
    try {
        if (!v.empty() && v[0] == '-') throw std::invalid_argument("negative");
        x = std::stoull(v, &pos, 0);
    } catch (std::exception&) {
        pos = 0;
    }
    if (pos != v.size() || v.empty())
        throw std::invalid_argument("the argument ('" + v + "') for option '--" + name + "' is invalid");
    return x;
}

inline double ParseDouble(const std::string& name, const std::string& v) {
    size_t pos = 0;
    double x = 0;
    try {
        x = std::stod(v, &pos);
    } catch (std::exception&) {
        pos = 0;
    }
    if (pos != v.size() || v.empty())
        throw std::invalid_argument("the argument ('" + v + "') for option '--" + name + "' is invalid");
    return x;
}

// Throws std::invalid_argument / std::runtime_error with the reference's messages; sets *help.
inline Common::Configuration Parse(int argc, char** argv, bool* help) {
    static const std::map<std::string, std::string> kShort = {
        {"-h", "help"}, {"-u", "unit"}, {"-o", "output"}, {"-f", "filename"}, {"-p", "partitions"}};
    static const std::set<std::string> kValued = {
        "primary", "secondary", "skew", "log", "join", "format", "unit", "output", "filename", "partitions",
        "hash", "seed", "table-seed", "data-seed", "radix-bits", "device", "gpus", "repeat", "table"};
    static const std::set<std::string> kFlags = {"help", "no-tma-store", "materialize", "stream-upload"};

    Common::Configuration c{};
    std::map<std::string, std::string> seen;
    *help = false;
    for (int i = 1; i < argc; ++i) {
        std::string a = argv[i], name, value;
        bool has_value = false;
        if (a.rfind("--", 0) == 0) {
            name = a.substr(2);
            auto eq = name.find('=');
            if (eq != std::string::npos) {
                value = name.substr(eq + 1);
                name = name.substr(0, eq);
                has_value = true;
            }
        } else if (kShort.count(a)) {
            name = kShort.at(a);
        } else {
            throw std::invalid_argument("unrecognised option '" + a + "'");
        }
        if (kFlags.count(name)) {
            seen[name] = "1";
            continue;
        }
        if (!kValued.count(name)) throw std::invalid_argument("unrecognised option '--" + name + "'");
        if (!has_value) {
            if (i + 1 >= argc) throw std::invalid_argument("the required argument for option '--" + name + "' is missing");
            value = argv[++i];
        }
        if (seen.count(name)) throw std::invalid_argument("option '--" + name + "' cannot be specified more than once");
        seen[name] = value;
    }
    if (seen.count("help")) {
        *help = true;
        return c;
    }
    auto get = [&](const char* k) -> const std::string* { auto it = seen.find(k); return it == seen.end() ? nullptr : &it->second; };
    if (auto v = get("primary")) c.PrimaryRelationSize = ParseUnsigned("primary", *v);
    if (auto v = get("secondary")) c.SecondaryRelationSize = ParseUnsigned("secondary", *v);
    if (auto v = get("skew")) c.SkewParameter = ParseDouble("skew", *v);
    if (auto v = get("log")) c.LoggerConfig.LogLevel = Common::SeverityLevelFromString(*v);
    if (auto v = get("format")) c.OutputFormatConfig.Format = Common::GetResultsFormatFromString(*v);
    if (auto v = get("unit")) c.OutputFormatConfig.TimeUnit = *v;
    if (auto v = get("output")) c.OutputConfig.Type = Common::GetOutputTypeFromString(*v);
    if (auto v = get("filename")) c.OutputConfig.File.Name = *v;
    if (auto v = get("partitions")) c.RadixClusteringConfig.NumberOfPartitions = ParseUnsigned("partitions", *v);
    if (auto v = get("hash")) c.Gpu.Hash = Common::GetHashFunctionFromString(*v);
    if (auto v = get("seed")) c.Gpu.Seed = ParseUnsigned("seed", *v);
    if (auto v = get("table-seed")) c.Gpu.TableSeed = ParseUnsigned("table-seed", *v);
    if (auto v = get("data-seed")) c.Gpu.DataSeed = static_cast<int64_t>(ParseUnsigned("data-seed", *v));
    if (auto v = get("device")) c.Gpu.Device = static_cast<int>(ParseUnsigned("device", *v));
    if (auto v = get("gpus")) c.Gpu.Gpus = static_cast<int>(ParseUnsigned("gpus", *v));
    if (auto v = get("table")) {
        if (*v == "separate-chaining") c.Gpu.ChainedTable = true;
        else if (*v != "linear-probing") throw std::invalid_argument("Unrecognized hash table type: " + *v + ".");
    }
    if (auto v = get("repeat")) c.Gpu.Repeat = static_cast<int>(ParseUnsigned("repeat", *v));
    if (auto v = get("radix-bits")) {
        auto comma = v->find(',');
        if (comma == std::string::npos) throw std::invalid_argument("--radix-bits expects b1,b2");
        c.Gpu.RadixBits[0] = static_cast<unsigned>(ParseUnsigned("radix-bits", v->substr(0, comma)));
        c.Gpu.RadixBits[1] = static_cast<unsigned>(ParseUnsigned("radix-bits", v->substr(comma + 1)));
    }
    c.Gpu.NoTmaStore = seen.count("no-tma-store") != 0;
    c.Gpu.Materialize = seen.count("materialize") != 0;
    c.Gpu.StreamUpload = seen.count("stream-upload") != 0;
    // --join is required (reference src/main.cpp:162-165)
    if (!get("join")) throw std::invalid_argument("the option '--join' is required but missing");
    c.JoinType = Common::GetJoinAlgorithmTypeFromString(*get("join"));

    // validateParsedConfiguration (reference src/Arguments.hpp:7-19)
    c.OutputConfig.Validate();
    c.OutputFormatConfig.Validate();
    if (c.JoinType != Common::JoinAlgorithmType::RadixParitioning && get("partitions"))
        throw std::invalid_argument(
            "validateParsedConfiguration: number of partitions can be specified only for RadixParitioning.");
    if (c.Gpu.Materialize && c.JoinType != Common::JoinAlgorithmType::RadixParitioning)
        throw std::invalid_argument("--materialize: the joined table is produced by the RadixParitioning joiner.");
    if (c.Gpu.StreamUpload && (c.Gpu.Materialize || c.Gpu.Repeat > 1))
        throw std::invalid_argument("--stream-upload joins once and count-only: not with --materialize or --repeat.");
    if (c.Gpu.Gpus < 1) throw std::invalid_argument("--gpus must be at least 1");
    if (c.Gpu.Gpus > 1 && c.Gpu.Materialize)
        throw std::invalid_argument("--gpus > 1 counts only: --materialize needs one GPU.");
    // The reference's default of 32 partitions is not GPUs x partitions per GPU for every GPU count (3, 5, 6, 7):
    // without -p the sharded join takes the library's own default, the largest multiple of the GPU count <= 64.
    if (c.Gpu.Gpus > 1 && c.Gpu.Gpus <= 64 && c.JoinType == Common::JoinAlgorithmType::RadixParitioning && !get("partitions"))
        c.RadixClusteringConfig.NumberOfPartitions = 64 / static_cast<size_t>(c.Gpu.Gpus) * static_cast<size_t>(c.Gpu.Gpus);
    return c;
}

}  // namespace Arguments
