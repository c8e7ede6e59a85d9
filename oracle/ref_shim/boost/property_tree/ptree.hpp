// Test-infrastructure shim for <boost/property_tree/ptree.hpp> (absent from this image) so that the
// reference's src/Common/Results.hpp:3-4,266-279 compiles unmodified. Keeps insertion order and
// dotted paths, which is all JSONResultsFormatter uses.
#pragma once
#include <sstream>
#include <string>
#include <utility>
#include <vector>

namespace boost {
namespace property_tree {
class ptree {
   public:
    template <typename T>
    void add(const std::string& path, const T& value) {
        std::ostringstream s;
        s << value;
        m_entries.emplace_back(path, s.str());
    }
    const std::vector<std::pair<std::string, std::string>>& entries() const { return m_entries; }

   private:
    std::vector<std::pair<std::string, std::string>> m_entries;
};
}  // namespace property_tree
}  // namespace boost
