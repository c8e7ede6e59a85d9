// phj_datagen.cpp -- host-side input generators of include/phj.h (phj_fill_sequential,
// phj_fill_zipf).
//
// Replaces DataGenerator::Sequential / DataGenerator::Zipf and the Park-Miller LCG they draw from
// (reference src/DataGenerator/Sequential.cpp:6-40, src/DataGenerator/Zipf.cpp:14-108,
// src/Common/Random.cpp:9-30). The generators produce the join's input; they are not on the hot
// path and stay on the host on purpose: Zipf::generate calls glibc pow(), whose results a device
// pow() does not reproduce bit for bit, and the north star wants the reference's generator output
// "with the same seed, then uploaded". The only change is that the seed is explicit: the reference
// seeds each batch from std::random_device (Zipf.cpp:86).
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <string>
#include <thread>
#include <vector>

#include "../../include/phj.h"

namespace {

// Minimal-standard multiplicative LCG (a = 16807, m = 2^31 - 1) evaluated with Schrage's trick,
// state held in a 64-bit signed integer like the reference's `long` on LP64.
struct Lcg {
    int64_t state;
    double next() {
        constexpr int64_t a = 16807, m = 2147483647, q = 127773, r = 2836;
        int64_t t = a * (state % q) - r * (state / q);
        state = t > 0 ? t : t + m;
        return static_cast<double>(state) / static_cast<double>(m);
    }
};

// Rejection-inversion Zipf sampler over [1, cardinality].
struct ZipfSampler {
    double alpha, sd, norm;
    ZipfSampler(double a, uint64_t cardinality) {
        alpha = a;
        sd = 1.001 - alpha;
        const double diff = 1.0 - alpha;
        if (std::abs(diff) < 0.01) {
            sd = 0.01 * (diff < 0 ? 1 : -1);
            alpha = 1.0 - sd;
        }
        norm = (std::pow(static_cast<double>(cardinality), sd) - alpha) / sd;
    }
    uint64_t draw(Lcg& g) const {
        for (;;) {
            const double u1 = g.next();
            const double u2 = g.next();
            const double x = u1 * norm;
            const double inv = x <= 1.0 ? x : std::pow(x * sd + alpha, 1.0 / sd);
            const double sample = std::floor(inv + 1);
            const double p_target = std::pow(sample, -alpha);
            const double p_proposal = sample <= 1.0 ? 1.0 / norm : std::pow(inv, -alpha) / norm;
            if (u2 < p_target / (p_proposal * norm)) return static_cast<uint64_t>(sample);
        }
    }
};

int worker_count(int32_t threads) {
    if (threads > 0) return threads;
    unsigned hc = std::thread::hardware_concurrency();
    return hc ? static_cast<int>(hc) : 1;
}

template <typename F>
void parallel_for_jobs(size_t jobs, int threads, F&& f) {
    threads = static_cast<int>(std::min<size_t>(static_cast<size_t>(threads), std::max<size_t>(jobs, 1)));
    if (threads <= 1) {
        for (size_t j = 0; j < jobs; ++j) f(j);
        return;
    }
    std::vector<std::thread> pool;
    for (int t = 0; t < threads; ++t)
        pool.emplace_back([&, t]() {
            for (size_t j = static_cast<size_t>(t); j < jobs; j += static_cast<size_t>(threads)) f(j);
        });
    for (auto& th : pool) th.join();
}

}  // namespace

extern "C" {

int phj_fill_sequential(phj_tuple* out, size_t n, int64_t start, int32_t threads) {
    if (n && !out) return PHJ_ERR_INVALID;
    const size_t chunk = 1 << 20;
    const size_t jobs = (n + chunk - 1) / chunk;
    parallel_for_jobs(jobs, worker_count(threads), [&](size_t j) {
        const size_t a = j * chunk, b = std::min(n, a + chunk);
        for (size_t i = a; i < b; ++i) {
            out[i].id = start + static_cast<int64_t>(i);
            out[i].payload = static_cast<int64_t>(i);
        }
    });
    return PHJ_OK;
}

int phj_fill_zipf(phj_tuple* out, size_t n, double alpha, int64_t range_first, int64_t range_second,
                  int64_t base_seed, size_t batches, int32_t threads) {
    if (n && !out) return PHJ_ERR_INVALID;
    if (alpha < 0.01) return PHJ_ERR_INVALID;               // Zipf.cpp:18-20
    if (range_first >= range_second) return PHJ_ERR_INVALID; // Zipf.cpp:61-67
    if (batches == 0) return PHJ_ERR_INVALID;
    // A seed congruent to 0 mod 2^31 - 1 pins the LCG at 0 and the sampler never terminates.
    for (size_t b = 0; b < batches; ++b)
        if ((base_seed + static_cast<int64_t>(b)) % 2147483647 == 0) return PHJ_ERR_INVALID;
    const uint64_t cardinality = static_cast<uint64_t>(range_second - range_first + 1);
    const int64_t correction = range_first - 1;
    const ZipfSampler sampler(alpha, cardinality);
    const size_t batch = n / batches;
    parallel_for_jobs(batches, worker_count(threads), [&](size_t b) {
        const size_t a = batch * b, e = (b + 1 == batches) ? n : batch * (b + 1);
        Lcg g{base_seed + static_cast<int64_t>(b)};
        for (size_t i = a; i < e; ++i) {
            out[i].id = static_cast<int64_t>(sampler.draw(g)) + correction;
            out[i].payload = static_cast<int64_t>(i);
        }
    });
    return PHJ_OK;
}

}  // extern "C"
