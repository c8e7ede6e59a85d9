// Timing sink and result rendering of the drop-in surface (reference src/Common/Results.hpp):
// IHashJoinTimer with its begin/end and set-duration halves, the timing result with its parameter
// map, and the JSON file output. The JSON is byte-compatible with what the reference writes through
// boost::property_tree::write_json (4-space indent, every leaf a quoted string, "parameters" in
// key order, then "results": partition / build / probe).
#pragma once
#include <chrono>
#include <fstream>
#include <map>
#include <memory>
#include <ostream>
#include <sstream>
#include <stdexcept>
#include <string>

#include "Configuration.hpp"

namespace Common {

class Parameters {
   public:
    using StorageType = std::map<std::string, std::string>;
    void SetParameter(const std::string& key, const std::string& value) { m_values[key] = value; }
    StorageType::const_iterator begin() const { return m_values.begin(); }
    StorageType::const_iterator end() const { return m_values.end(); }

   private:
    StorageType m_values;
};

class HashJoinTimingResult {
   public:
    HashJoinTimingResult() = default;
    HashJoinTimingResult(std::chrono::nanoseconds build, std::chrono::nanoseconds probe,
                         std::chrono::nanoseconds partitioning, const Parameters& parameters)
        : m_parameters(parameters), m_build(build), m_probe(probe), m_partitioning(partitioning) {}

    void SetBuildPhaseDuration(std::chrono::nanoseconds d) { m_build = d; }
    void SetProbePhaseDuration(std::chrono::nanoseconds d) { m_probe = d; }
    void SetPartitioningPhaseDuration(std::chrono::nanoseconds d) { m_partitioning = d; }
    void SetParameters(const Parameters& p) { m_parameters = p; }
    std::chrono::nanoseconds GetBuildPhaseDuration() const { return m_build; }
    std::chrono::nanoseconds GetProbePhaseDuration() const { return m_probe; }
    std::chrono::nanoseconds GetPartitioningPhaseDuration() const { return m_partitioning; }
    const Parameters& GetParameters() const { return m_parameters; }

   private:
    Parameters m_parameters;
    std::chrono::nanoseconds m_build{0}, m_probe{0}, m_partitioning{0};
};

// Same virtual surface as the reference (src/Common/Results.hpp:131-149). The GPU joiners report
// device-measured durations, so they drive the Set*Duration half.
class IHashJoinTimer {
   public:
    virtual void SetBuildPhaseBegin() = 0;
    virtual void SetBuildPhaseEnd() = 0;
    virtual void SetPartitioningPhaseBegin() = 0;
    virtual void SetPartitioningPhaseEnd() = 0;
    virtual void SetProbePhaseBegin() = 0;
    virtual void SetProbePhaseEnd() = 0;
    virtual void SetBuildPhaseDuration(std::chrono::nanoseconds duration) = 0;
    virtual void SetProbePhaseDuration(std::chrono::nanoseconds duration) = 0;
    virtual void SetPartitionPhaseDuration(std::chrono::nanoseconds duration) = 0;
    virtual HashJoinTimingResult GetResult() = 0;
    virtual ~IHashJoinTimer() = default;
};

class NoOpHashJoinTimer final : public IHashJoinTimer {
   public:
    void SetBuildPhaseBegin() override {}
    void SetBuildPhaseEnd() override {}
    void SetPartitioningPhaseBegin() override {}
    void SetPartitioningPhaseEnd() override {}
    void SetProbePhaseBegin() override {}
    void SetProbePhaseEnd() override {}
    void SetBuildPhaseDuration(std::chrono::nanoseconds) override {}
    void SetProbePhaseDuration(std::chrono::nanoseconds) override {}
    void SetPartitionPhaseDuration(std::chrono::nanoseconds) override {}
    HashJoinTimingResult GetResult() override { return HashJoinTimingResult(); }
};

class HashJoinTimer final : public IHashJoinTimer {
    using Clock = std::chrono::steady_clock;

   public:
    explicit HashJoinTimer(const Parameters& parameters) : m_parameters(parameters) {}

    void SetBuildPhaseBegin() override { m_buildStart = Clock::now(); }
    void SetBuildPhaseEnd() override { m_build = Clock::now() - m_buildStart; }
    void SetPartitioningPhaseBegin() override { m_partitioningStart = Clock::now(); }
    void SetPartitioningPhaseEnd() override { m_partitioning = Clock::now() - m_partitioningStart; }
    void SetProbePhaseBegin() override { m_probeStart = Clock::now(); }
    // Unlike the reference (src/Common/Results.hpp:202 subtracts the *build* start, so its NPJ
    // "probe" contains the build), the probe phase is measured from its own start.
    void SetProbePhaseEnd() override { m_probe = Clock::now() - m_probeStart; }

    void SetBuildPhaseDuration(std::chrono::nanoseconds d) override { m_build = d; }
    void SetProbePhaseDuration(std::chrono::nanoseconds d) override { m_probe = d; }
    void SetPartitionPhaseDuration(std::chrono::nanoseconds d) override { m_partitioning = d; }

    HashJoinTimingResult GetResult() override {
        return HashJoinTimingResult(m_build, m_probe, m_partitioning, m_parameters);
    }

   private:
    Parameters m_parameters;
    std::chrono::nanoseconds m_build{0}, m_probe{0}, m_partitioning{0};
    Clock::time_point m_buildStart, m_probeStart, m_partitioningStart;
};

class IResultsFormatter {
   public:
    virtual void Format(std::ostream& stream, const HashJoinTimingResult& result) = 0;
    virtual ~IResultsFormatter() = default;
};

class IResultsRenderer {
   public:
    virtual void Render(std::shared_ptr<IResultsFormatter> formatter, const HashJoinTimingResult& result) = 0;
    virtual ~IResultsRenderer() = default;
};

class JSONResultsFormatter final : public IResultsFormatter {
   public:
    explicit JSONResultsFormatter(const ResultsFormatConfiguration& config) : m_config(config) {}

    void Format(std::ostream& os, const HashJoinTimingResult& r) override {
        os << "{\n    \"id\": \"hashjointimingresult\",\n    \"parameters\": {\n";
        size_t i = 0, n = 0;
        for (auto it = r.GetParameters().begin(); it != r.GetParameters().end(); ++it) ++n;
        for (const auto& kv : r.GetParameters())
            os << "        \"" << kv.first << "\": \"" << kv.second << "\"" << (++i == n ? "\n" : ",\n");
        os << "    },\n    \"results\": {\n"
           << "        \"partition\": \"" << Cast(r.GetPartitioningPhaseDuration()) << "\",\n"
           << "        \"build\": \"" << Cast(r.GetBuildPhaseDuration()) << "\",\n"
           << "        \"probe\": \"" << Cast(r.GetProbePhaseDuration()) << "\"\n"
           << "    }\n}\n";
    }

   private:
    // Truncating conversion, as std::chrono::duration_cast does in the reference (:284-301).
    std::string Cast(std::chrono::nanoseconds d) const {
        long long v = d.count();
        if (m_config.TimeUnit == "ns") {
        } else if (m_config.TimeUnit == "us") {
            v /= 1000;
        } else if (m_config.TimeUnit == "ms") {
            v /= 1000000;
        } else if (m_config.TimeUnit == "s") {
            v /= 1000000000;
        } else {
            throw std::runtime_error("JSONResultsFormatter::CastDurationToString: unrecognized duration unit: " +
                                     m_config.TimeUnit);
        }
        return std::to_string(v);
    }
    const ResultsFormatConfiguration m_config;
};

class FileResultsRenderer final : public IResultsRenderer {
   public:
    explicit FileResultsRenderer(const OutputConfiguration& config) : m_file(config.File.Name) {}
    void Render(std::shared_ptr<IResultsFormatter> formatter, const HashJoinTimingResult& result) override {
        formatter->Format(m_file, result);
        m_file.flush();
    }

   private:
    std::ofstream m_file;
};

inline std::shared_ptr<IResultsFormatter> SelectResultsFormatter(const Configuration& config) {
    return std::make_shared<JSONResultsFormatter>(config.OutputFormatConfig);
}
inline std::shared_ptr<IResultsRenderer> SelectResultsRenderer(const Configuration& config) {
    return std::make_shared<FileResultsRenderer>(config.OutputConfig);
}

}  // namespace Common
