// Test-infrastructure shim for <boost/property_tree/json_parser.hpp>: write_json with Boost's
// layout (4-space indent, every leaf a quoted string), enough to reproduce files such as
// results/1.05/partitions_32.txt from the reference's own JSONResultsFormatter.
#pragma once
#include <map>
#include <ostream>
#include <string>
#include <vector>

#include "ptree.hpp"

namespace boost {
namespace property_tree {
namespace json_parser {
inline void write_json(std::ostream& os, const ptree& pt) {
    // Group "a.b" paths by their first component, preserving first-seen order.
    std::vector<std::string> order;
    std::map<std::string, std::vector<std::pair<std::string, std::string>>> groups;
    for (const auto& e : pt.entries()) {
        auto dot = e.first.find('.');
        std::string head = dot == std::string::npos ? e.first : e.first.substr(0, dot);
        std::string tail = dot == std::string::npos ? std::string() : e.first.substr(dot + 1);
        if (!groups.count(head)) order.push_back(head);
        groups[head].emplace_back(tail, e.second);
    }
    os << "{\n";
    for (size_t i = 0; i != order.size(); ++i) {
        const auto& g = groups[order[i]];
        if (g.size() == 1 && g[0].first.empty()) {
            os << "    \"" << order[i] << "\": \"" << g[0].second << "\"";
        } else {
            os << "    \"" << order[i] << "\": {\n";
            for (size_t j = 0; j != g.size(); ++j) {
                os << "        \"" << g[j].first << "\": \"" << g[j].second << "\""
                   << (j + 1 == g.size() ? "\n" : ",\n");
            }
            os << "    }";
        }
        os << (i + 1 == order.size() ? "\n" : ",\n");
    }
    os << "}\n";
}
}  // namespace json_parser
using json_parser::write_json;
}  // namespace property_tree
}  // namespace boost
