// Host data model of the drop-in surface: the layouts the reference's joiners consume and return
// (reference src/Common/Table.hpp:20-57). Tuple is the 16-byte record the device kernels read in
// place; Table<T> is a named, contiguous, 16-byte-aligned array of records.
#pragma once
#include <cstddef>
#include <cstdint>
#include <ostream>
#include <string>
#include <vector>

namespace Common {

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
    std::vector<TupleType> m_tuples;
};

}  // namespace Common
