// Test-infrastructure shim (NOT product code, NOT reference code).
// Stands in for <boost/log/sources/severity_logger.hpp>, which is absent from this image, so that
// the reference's hot-path headers (src/Common/Logger.hpp:3-5,16,33) compile UNMODIFIED into
// oracle/_ref/libphj_ref.so. Every log line is handed to phj_shim::sink(), which is how the
// harness reads the "Joined N tuples" record -- the only channel on which the reference reports
// its match count (src/NoPartitioning/HashJoin.hpp:184, src/RadixCluster/HashJoin.hpp:320-321).
#pragma once
#include <sstream>
#include <string>

namespace phj_shim {
void sink(int severity, const std::string& line);

class LineCapture {
   public:
    explicit LineCapture(int severity) : m_severity(severity) {}
    ~LineCapture() { sink(m_severity, m_stream.str()); }
    std::ostream& stream() { return m_stream; }

   private:
    int m_severity;
    std::ostringstream m_stream;
};
}  // namespace phj_shim

namespace boost {
namespace log {
namespace sources {
template <typename LevelT>
class severity_logger {
   public:
    template <typename A, typename B>
    void add_attribute(const A&, const B&) {}
};
}  // namespace sources
}  // namespace log
}  // namespace boost

#define BOOST_LOG_SEV(lg, sev) ::phj_shim::LineCapture(static_cast<int>(sev)).stream()
