// fake_nccl.cpp -- TEST INFRASTRUCTURE (tests/emu): the eight NCCL entry points phj_dist.inl binds with dlopen
// ("libnccl.so.2"), for ranks that are host threads of ONE process over host memory. Built as
// tests/emu/_build/libnccl.so.2 and found through LD_LIBRARY_PATH by the emulated test runs only.
// A collective blocks its calling thread until every rank of the communicator has called it (the fake CUDA
// runtime executes stream work at once, so "enqueued" means "done").
#include <nccl.h>
#include <stdint.h>
#include <string.h>

#include <condition_variable>
#include <map>
#include <mutex>
#include <vector>

namespace {

struct Group {
    int n = 0;
    std::mutex m;
    std::condition_variable cv;
    int arrived = 0;
    uint64_t gen = 0;
    const void* send[64] = {};
    int joined = 0, left = 0;
    void barrier() {
        std::unique_lock<std::mutex> lk(m);
        const uint64_t g = gen;
        if (++arrived == n) {
            arrived = 0;
            ++gen;
            cv.notify_all();
        } else {
            cv.wait(lk, [&] { return gen != g; });
        }
    }
};

struct Comm {
    Group* group;
    int rank;
};

std::mutex g_mutex;
std::map<uint64_t, Group*> g_pending;  // unique id -> group still collecting its ranks
uint64_t g_next_id = 1;

size_t type_size(ncclDataType_t t) {
    switch (t) {
        case ncclInt8: case ncclUint8: return 1;
        case ncclFloat16: return 2;
        case ncclInt32: case ncclUint32: case ncclFloat32: return 4;
        default: return 8;
    }
}

void release(Comm* c) {
    bool last;
    {
        std::lock_guard<std::mutex> lk(g_mutex);
        last = ++c->group->left == c->group->n;
    }
    if (last) delete c->group;
    delete c;
}

}  // namespace

extern "C" {

ncclResult_t ncclGetVersion(int* v) {
    *v = 0;
    return ncclSuccess;
}
const char* ncclGetErrorString(ncclResult_t) { return "emulated NCCL error"; }

ncclResult_t ncclGetUniqueId(ncclUniqueId* id) {
    std::lock_guard<std::mutex> lk(g_mutex);
    memset(id, 0, sizeof(*id));
    const uint64_t v = g_next_id++;
    memcpy(id, &v, sizeof(v));
    return ncclSuccess;
}

// Ranks of one communicator are threads of this process that present the same id.
ncclResult_t ncclCommInitRank(ncclComm_t* comm, int nranks, ncclUniqueId id, int rank) {
    uint64_t key;
    memcpy(&key, &id, sizeof(key));
    Group* g;
    {
        std::lock_guard<std::mutex> lk(g_mutex);
        Group*& slot = g_pending[key];
        if (!slot) {
            slot = new Group;
            slot->n = nranks;
        }
        g = slot;
        if (++g->joined == nranks) g_pending.erase(key);
    }
    if (g->n != nranks || rank < 0 || rank >= nranks) return ncclInvalidArgument;
    *comm = reinterpret_cast<ncclComm_t>(new Comm{g, rank});
    g->barrier();  // like the real call: returns once every rank has joined
    return ncclSuccess;
}

ncclResult_t ncclCommInitAll(ncclComm_t* comms, int ndev, const int*) {
    Group* g = new Group;
    g->n = ndev;
    g->joined = ndev;
    for (int r = 0; r < ndev; ++r) comms[r] = reinterpret_cast<ncclComm_t>(new Comm{g, r});
    return ncclSuccess;
}

ncclResult_t ncclCommDestroy(ncclComm_t comm) {
    release(reinterpret_cast<Comm*>(comm));
    return ncclSuccess;
}

ncclResult_t ncclAllGather(const void* send, void* recv, size_t count, ncclDataType_t type, ncclComm_t comm,
                           cudaStream_t) {
    Comm* c = reinterpret_cast<Comm*>(comm);
    Group* g = c->group;
    const size_t bytes = count * type_size(type);
    g->send[c->rank] = send;
    g->barrier();  // every rank's contribution is known
    for (int r = 0; r < g->n; ++r) {
        char* dst = static_cast<char*>(recv) + (size_t)r * bytes;
        if (dst != g->send[r]) memmove(dst, g->send[r], bytes);
    }
    g->barrier();  // ... and has been read by everybody
    return ncclSuccess;
}

ncclResult_t ncclAllReduce(const void* send, void* recv, size_t count, ncclDataType_t type, ncclRedOp_t op,
                           ncclComm_t comm, cudaStream_t) {
    if (type != ncclUint64 || op != ncclSum) return ncclInvalidArgument;  // all the engine uses
    Comm* c = reinterpret_cast<Comm*>(comm);
    Group* g = c->group;
    g->send[c->rank] = send;
    g->barrier();
    std::vector<uint64_t> sum(count, 0);
    for (int r = 0; r < g->n; ++r)
        for (size_t i = 0; i < count; ++i) sum[i] += static_cast<const uint64_t*>(g->send[r])[i];
    g->barrier();  // everybody has read every contribution: in-place results may be written
    memcpy(recv, sum.data(), count * 8);
    return ncclSuccess;
}

}  // extern "C"
