// phjoin -- the reference's benchmark driver on the B200 engine.
//
// Same flow as the reference's main (src/main.cpp:210-293): parse arguments, generate the primary
// (Sequential) and secondary (Zipf) relations, run the selected join once through the joiner
// interface, render the timing result as JSON to the output file. On any parse/validation error it
// prints the message and the option table and exits with 1, like the reference (:199-205); a join
// failure is logged and exits with 1 (:277-281).
#include <chrono>
#include <cstdio>
#include <iostream>
#include <memory>
#include <string>

#include "Arguments.hpp"
#include "Common/Configuration.hpp"
#include "Common/Results.hpp"
#include "Common/Table.hpp"
#include "DataGenerator/Generators.hpp"
#include "Gpu/HashJoin.hpp"

namespace Common {
std::string generate_uuid() {
    static unsigned long long counter = 0;
    char buf[64];
    std::snprintf(buf, sizeof(buf), "phj-table-%llu", counter++);
    return buf;
}
}  // namespace Common

namespace {
struct Log {
    Common::SeverityLevel level;
    void operator()(Common::SeverityLevel sev, const std::string& msg) const {
        static const char* names[] = {"trace", "debug", "info", "error", "critical"};
        if (sev >= level) std::clog << "main(" << names[sev] << ") " << msg << "\n";
    }
};
}  // namespace

int main(int argc, char** argv) {
    Common::Configuration config;
    try {
        bool help = false;
        config = Arguments::Parse(argc, argv, &help);
        if (help) {
            std::cout << Arguments::HelpText() << "\n";
            return 0;
        }
    } catch (std::exception& e) {
        std::cout << e.what() << "\n\n" << Arguments::HelpText() << "\n";
        return 1;
    }
    Log log{config.LoggerConfig.LogLevel};
    auto formatter = Common::SelectResultsFormatter(config);
    auto renderer = Common::SelectResultsRenderer(config);
    log(Common::info, "Starting running tests.");

    // generateTables (reference src/main.cpp:35-79): primary = Sequential from 1, secondary = Zipf
    // over [1, |primary|]
    log(Common::debug, "Generating primary relation with size " + std::to_string(config.PrimaryRelationSize) +
                           " and secondary relation with size " + std::to_string(config.SecondaryRelationSize) + ".");
    auto primary = std::make_shared<Common::Table<Common::Tuple>>(config.PrimaryRelationSize, Common::generate_uuid());
    auto secondary = std::make_shared<Common::Table<Common::Tuple>>(config.SecondaryRelationSize, Common::generate_uuid());
    Common::HashJoinTimingResult result;
    uint64_t matches = 0, total_ns = 0, alg_bytes = 0, e2e_ns = 0, chunks = 0;
    try {
        DataGenerator::Sequential::FillTable(primary, {1});
        DataGenerator::Zipf::FillTable(secondary, {config.SkewParameter, {1, config.PrimaryRelationSize},
                                                   config.Gpu.DataSeed, config.Gpu.DataBatches});
        log(Common::debug, "Generation of relations finished.");

        Common::Parameters params;
        params.SetParameter("PrimaryRelationSize", std::to_string(config.PrimaryRelationSize));
        params.SetParameter("SecondaryRelationSize", std::to_string(config.SecondaryRelationSize));
        params.SetParameter("Skew", std::to_string(config.SkewParameter));
        std::shared_ptr<Common::IHashJoinTimer> timer;
        if (config.JoinType == Common::JoinAlgorithmType::NoPartitioning) {
            log(Common::debug, "Executing NoPartitionHashJoin algorithm.");
            params.SetParameter("Type", "NoPartitioning");
            timer = std::make_shared<Common::HashJoinTimer>(params);
            Gpu::NoPartitioning::HashJoiner joiner(config.NoPartitioningConfig, config.Gpu);
            joiner.Run(primary, secondary, timer);
            matches = joiner.GetLastResult().matches;
            total_ns = joiner.GetLastResult().total_ns;
            alg_bytes = joiner.GetLastResult().hbm_bytes_alg;
            e2e_ns = joiner.GetLastResult().e2e_ns;
            chunks = joiner.GetLastResult().upload_chunks;
        } else {
            log(Common::debug, "Executing Radix Clustering join algorithm.");
            params.SetParameter("Type", "RadixParitioning");  // sic, as in the reference's JSON
            params.SetParameter("NumberOfPartitions", std::to_string(config.RadixClusteringConfig.NumberOfPartitions));
            timer = std::make_shared<Common::HashJoinTimer>(params);
            Gpu::RadixClustering::HashJoiner joiner(config.RadixClusteringConfig, config.Gpu);
            auto joined = joiner.Run(primary, secondary, timer);
            matches = joiner.GetLastResult().matches;
            total_ns = joiner.GetLastResult().total_ns;
            alg_bytes = joiner.GetLastResult().hbm_bytes_alg;
            e2e_ns = joiner.GetLastResult().e2e_ns;
            chunks = joiner.GetLastResult().upload_chunks;
            if (config.Gpu.Materialize)
                log(Common::info, "Joined table holds " + std::to_string(joined->GetSize()) + " rows.");
        }
        result = timer->GetResult();
    } catch (std::exception& e) {
        log(Common::error, std::string("Hash join algorithm stopped due to exception begin raised: ") + e.what());
        return 1;
    }
    log(Common::debug, "Joined " + std::to_string(matches) + " tuples.");
    if (total_ns) {
        const double tuples = static_cast<double>(config.PrimaryRelationSize + config.SecondaryRelationSize);
        char buf[256];
        std::snprintf(buf, sizeof(buf), "Device time %.3f ms, %.2f G tuples/s, %.0f GB/s of algorithmic HBM traffic.",
                      total_ns / 1e6, tuples / total_ns, static_cast<double>(alg_bytes) / total_ns);
        log(Common::info, buf);
    }
    if (e2e_ns) {
        char buf[256];
        std::snprintf(buf, sizeof(buf), "End to end (upload in %llu chunk(s) + join) %.3f ms.",
                      static_cast<unsigned long long>(chunks), e2e_ns / 1e6);
        log(Common::info, buf);
    }
    renderer->Render(formatter, result);
    log(Common::info, "Finished running tests.");
    return 0;
}
