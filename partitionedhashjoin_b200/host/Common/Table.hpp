// Host data model of the drop-in surface: the layouts the reference's joiners consume and return
// (reference src/Common/Table.hpp:20-57). Tuple is the 16-byte record the device kernels read in
// place; Table<T> is a named, contiguous, 16-byte-aligned array of records.
#pragma once
#include <cstddef>
#include <cstdint>
#include <cstdlib>
#include <mutex>
#include <new>
#include <ostream>
#include <string>
#include <unordered_set>
#include <vector>

#include "../../../include/phj.h"

namespace Common {

// Relations live in PAGE-LOCKED host memory (phj_host_alloc): the joiners' Run(tableA, tableB) uploads them,
// and an upload from pinned memory runs at PCIe speed (measured 55 GB/s) where one from pageable memory -- what
// the reference's std::vector holds -- reaches 11 GB/s (bench.py: e2e vs e2e.pageable). Without a CUDA device
// (argument parsing, JSON rendering tests) the allocator falls back to ordinary aligned memory.
template <typename T>
struct PinnedAllocator {
    using value_type = T;
    PinnedAllocator() = default;
    template <typename U>
    PinnedAllocator(const PinnedAllocator<U>&) {}

    T* allocate(size_t n) {
        void* p = nullptr;
        if (phj_host_alloc(&p, n * sizeof(T)) == PHJ_OK && p) {
            std::lock_guard<std::mutex> lk(Mutex());
            Pinned().insert(p);
            return static_cast<T*>(p);
        }
        p = std::aligned_alloc(64, (n * sizeof(T) + 63) / 64 * 64);
        if (!p) throw std::bad_alloc();
        return static_cast<T*>(p);
    }
    void deallocate(T* p, size_t) {
        bool pinned;
        {
            std::lock_guard<std::mutex> lk(Mutex());
            pinned = Pinned().erase(p) != 0;
        }
        if (pinned) phj_host_free(p);
        else std::free(p);
    }
    template <typename U>
    bool operator==(const PinnedAllocator<U>&) const { return true; }
    template <typename U>
    bool operator!=(const PinnedAllocator<U>&) const { return false; }

   private:
    static std::mutex& Mutex() {
        static std::mutex m;
        return m;
    }
    static std::unordered_set<void*>& Pinned() {
        static std::unordered_set<void*> s;
        return s;
    }
};

struct alignas(16) Tuple {
    int64_t id;
    int64_t payload;
};

struct JoinedTuple {
    int64_t id;
    int64_t payloadA;
    int64_t payloadB;
};

inline std::ostream& operator<<(std::ostream& os, const Tuple& t) { return os << t.id << ", " << t.payload; }
inline std::ostream& operator<<(std::ostream& os, const JoinedTuple& t) {
    return os << t.id << ", " << t.payloadA << ", " << t.payloadB;
}

// Identifier for a table instance. The reference draws a random UUID (boost.uuid,
// src/Common/Table.cpp:18-22); uniqueness within the process is all any caller relies on.
std::string generate_uuid();

template <typename TupleType>
class Table {
   public:
    explicit Table(std::string id) : m_id(std::move(id)) {}
    Table(size_t size, std::string id) : m_id(std::move(id)), m_tuples(size) {}

    TupleType& operator[](size_t i) { return m_tuples[i]; }
    const TupleType& operator[](size_t i) const { return m_tuples[i]; }
    size_t GetSize() const { return m_tuples.size(); }
    size_t GetCapacity() const { return m_tuples.capacity(); }
    const std::string& GetID() const { return m_id; }

    void Resize(size_t size) { m_tuples.resize(size); }

    // What crosses the C ABI: &(*table)[0] and GetSize().
    TupleType* Data() { return m_tuples.data(); }
    const TupleType* Data() const { return m_tuples.data(); }

   private:
    std::string m_id;
    std::vector<TupleType, PinnedAllocator<TupleType>> m_tuples;
};

}  // namespace Common
