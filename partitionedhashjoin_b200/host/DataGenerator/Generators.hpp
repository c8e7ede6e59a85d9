// Input generators with the reference's shapes: DataGenerator::Sequential::FillTable and
// DataGenerator::Zipf::FillTable (reference src/DataGenerator/Sequential.hpp:10-21,
// src/DataGenerator/Zipf.hpp:13-29). They run on the host (phj_fill_* in include/phj.h) because the
// north star wants the reference generator's output uploaded, and return when the table is
// filled (the reference returns a future of its thread-pool tasks). The Zipf seeding is explicit:
// `batches` equal slices, slice b drawing from an LCG seeded `seed + b`.
#pragma once
#include <cstdint>
#include <memory>
#include <sstream>
#include <stdexcept>
#include <utility>

#include "../../../include/phj.h"
#include "../Common/Table.hpp"

namespace DataGenerator {

class Sequential {
   public:
    struct Parameters {
        int64_t start;
        int threads = 0;
    };
    static void FillTable(std::shared_ptr<Common::Table<Common::Tuple>> table, const Parameters& parameters) {
        if (phj_fill_sequential(reinterpret_cast<phj_tuple*>(table->Data()), table->GetSize(), parameters.start,
                                parameters.threads) != PHJ_OK)
            throw std::invalid_argument("Sequential::FillTable: invalid arguments");
    }
};

class Zipf {
   public:
    struct Parameters {
        double alpha;
        std::pair<size_t, size_t> range;
        int64_t seed = 12345;
        size_t batches = 64;
        int threads = 0;
    };
    static void FillTable(std::shared_ptr<Common::Table<Common::Tuple>> table, const Parameters& parameters) {
        if (parameters.range.first >= parameters.range.second) {
            std::ostringstream msg;  // same message as the reference (src/DataGenerator/Zipf.cpp:61-67)
            msg << "Range for Zipf generation is incorrectly specified: [" << parameters.range.first << ", "
                << parameters.range.second << "].";
            throw std::invalid_argument(msg.str());
        }
        if (parameters.alpha < 0.01)  // src/DataGenerator/Zipf.cpp:18-20
            throw std::invalid_argument("Skew parameter must be greater than 0.01.");
        if (phj_fill_zipf(reinterpret_cast<phj_tuple*>(table->Data()), table->GetSize(), parameters.alpha,
                          static_cast<int64_t>(parameters.range.first), static_cast<int64_t>(parameters.range.second),
                          parameters.seed, parameters.batches, parameters.threads) != PHJ_OK)
            throw std::invalid_argument("Zipf::FillTable: invalid seed or batch count");
    }
};

}  // namespace DataGenerator
