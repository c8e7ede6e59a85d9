// fake_cudart.cpp -- TEST INFRASTRUCTURE (tests/emu): the 41 CUDA runtime entry points the engine
// (partitionedhashjoin_b200/csrc/phj_engine.cu + phj_dist.inl, compiled UNMODIFIED by nvcc with -cudart none) calls,
// implemented on the host so that the engine's own orchestration -- plans, launch order, streams, events, the
// multi-GPU group with its host threads -- runs on a box without a GPU, with every kernel launch handed to the
// emulator (emu_core.cpp) that executes the g++-compiled kernel source. Never linked into the product library.
//
// Model: "device" memory is host memory; a peer GPU's memory is just another pointer; everything a stream is given
// happens at once on the calling thread (a legal execution of the stream program: the engine enqueues work in
// dependency order, and where another host thread is involved it waits for that thread's enqueue explicitly);
// events are time stamps; PHJ_EMU_GPUS devices (default 8) of PHJ_EMU_SMS SMs each (default 2: small grids).
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include <map>
#include <mutex>
#include <string>
#include <vector>

namespace {

using Tramp = void (*)(void**);
using FindFn = Tramp (*)(const char*);
using RunFn = int (*)(Tramp, void**, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned);

std::mutex g_mutex;
std::map<const void*, std::string> g_kernels;  // host stub -> mangled kernel name
std::map<const void*, Tramp> g_tramps;
FindFn g_find = nullptr;
RunFn g_run = nullptr;
thread_local int tl_device = 0;
thread_local cudaError_t tl_last_error = cudaSuccess;  // of a kernel launch: `<<<>>>` has no return value

struct CallConfig {
    dim3 grid, block;
    size_t smem;
    void* stream;
};
thread_local std::vector<CallConfig> tl_config;

struct FakeEvent {
    double ms = 0;
};

int env_int(const char* name, int fallback) {
    const char* v = getenv(name);
    return v && *v ? atoi(v) : fallback;
}

double now_ms() {
    timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6;
}

bool load_kernels() {
    if (g_run) return true;
    Dl_info info;
    if (!dladdr((const void*)&load_kernels, &info) || !info.dli_fname) return false;
    std::string path = info.dli_fname;
    path = path.substr(0, path.find_last_of('/') + 1) + "libphj_emu_kernels.so";
    void* lib = dlopen(path.c_str(), RTLD_NOW | RTLD_LOCAL);
    if (!lib) {
        fprintf(stderr, "emu: %s\n", dlerror());
        return false;
    }
    g_find = (FindFn)dlsym(lib, "emu_find_kernel");
    g_run = (RunFn)dlsym(lib, "emu_run_kernel");
    return g_find && g_run;
}

}  // namespace

extern "C" {

// ---- what nvcc's host stubs call --------------------------------------------------------------------
void** __cudaRegisterFatBinary(void*) {
    static void* handle = nullptr;
    return &handle;
}
void __cudaRegisterFatBinaryEnd(void**) {}
void __cudaUnregisterFatBinary(void**) {}
void __cudaRegisterFunction(void**, const char* host_fun, char*, const char* device_name, int, uint3*, uint3*, dim3*,
                            dim3*, int*) {
    std::lock_guard<std::mutex> lk(g_mutex);
    g_kernels[host_fun] = device_name;
}
unsigned __cudaPushCallConfiguration(dim3 grid, dim3 block, size_t smem, void* stream) {
    tl_config.push_back(CallConfig{grid, block, smem, stream});
    return 0;
}
cudaError_t __cudaPopCallConfiguration(dim3* grid, dim3* block, size_t* smem, void* stream) {
    if (tl_config.empty()) return cudaErrorInvalidConfiguration;
    const CallConfig c = tl_config.back();
    tl_config.pop_back();
    *grid = c.grid;
    *block = c.block;
    *smem = c.smem;
    *(void**)stream = c.stream;
    return cudaSuccess;
}

static cudaError_t launch_kernel(const void* func, dim3 grid, dim3 block, void** args, size_t smem);

cudaError_t cudaLaunchKernel(const void* func, dim3 grid, dim3 block, void** args, size_t smem, cudaStream_t) {
    const cudaError_t e = launch_kernel(func, grid, block, args, smem);
    if (e != cudaSuccess) {
        tl_last_error = e;  // what cudaGetLastError() reports, as on the device
        fprintf(stderr, "emu: kernel launch failed with error %d (grid %u x %u x %u, block %u, %zu bytes of shared memory)\n",
                (int)e, grid.x, grid.y, grid.z, block.x * block.y * block.z, smem);
    }
    return e;
}

static cudaError_t launch_kernel(const void* func, dim3 grid, dim3 block, void** args, size_t smem) {
    Tramp tramp = nullptr;
    {
        std::lock_guard<std::mutex> lk(g_mutex);
        if (!load_kernels()) return cudaErrorInitializationError;
        auto it = g_tramps.find(func);
        if (it == g_tramps.end()) {
            auto name = g_kernels.find(func);
            if (name == g_kernels.end()) return cudaErrorInvalidDeviceFunction;
            tramp = g_find(name->second.c_str());
            if (!tramp) {
                fprintf(stderr, "emu: kernel %s is not in the emulated build\n", name->second.c_str());
                return cudaErrorInvalidDeviceFunction;
            }
            g_tramps[func] = tramp;
        } else {
            tramp = it->second;
        }
    }
    if (getenv("PHJ_EMU_TRACE")) {
        std::lock_guard<std::mutex> lk(g_mutex);
        fprintf(stderr, "emu: launch %s <<<(%u,%u,%u), (%u,%u,%u), %zu>>>\n", g_kernels[func].c_str(), grid.x, grid.y, grid.z,
                block.x, block.y, block.z, smem);
    }
    if (smem > 232448) return cudaErrorInvalidValue;
    if (grid.x == 0 || grid.y == 0 || grid.z == 0) return cudaErrorInvalidConfiguration;
    return g_run(tramp, args, grid.x, grid.y, grid.z, block.x, block.y, block.z) == 0 ? cudaSuccess
                                                                                       : cudaErrorLaunchFailure;
}

// ---- devices ------------------------------------------------------------------------------------------
cudaError_t cudaGetDeviceCount(int* n) {
    *n = env_int("PHJ_EMU_GPUS", 8);
    return cudaSuccess;
}
cudaError_t cudaSetDevice(int d) {
    if (d < 0 || d >= env_int("PHJ_EMU_GPUS", 8)) return cudaErrorInvalidDevice;
    tl_device = d;
    return cudaSuccess;
}
cudaError_t cudaGetDeviceProperties_v2(cudaDeviceProp* p, int) {
    memset(p, 0, sizeof(*p));
    snprintf(p->name, sizeof(p->name), "emulated sm_100 (tests/emu: kernel source on the CPU)");
    p->major = 10;
    p->minor = 0;
    p->multiProcessorCount = env_int("PHJ_EMU_SMS", 2);
    p->sharedMemPerBlockOptin = 232448;
    p->sharedMemPerBlock = 49152;
    p->l2CacheSize = 126 << 20;
    p->totalGlobalMem = (size_t)32 << 30;
    p->warpSize = 32;
    p->maxThreadsPerBlock = 1024;
    return cudaSuccess;
}
cudaError_t cudaDeviceGetAttribute(int* v, cudaDeviceAttr, int) {
    *v = 1000;
    return cudaSuccess;
}
cudaError_t cudaDeviceGetStreamPriorityRange(int* lo, int* hi) {
    *lo = 0;
    *hi = -1;
    return cudaSuccess;
}
cudaError_t cudaDeviceSynchronize(void) { return cudaSuccess; }
cudaError_t cudaDeviceCanAccessPeer(int* can, int, int) {
    *can = 1;
    return cudaSuccess;
}
cudaError_t cudaDeviceEnablePeerAccess(int, unsigned) { return cudaSuccess; }
cudaError_t cudaFuncSetAttribute(const void*, cudaFuncAttribute, int) { return cudaSuccess; }
cudaError_t cudaGetLastError(void) {
    const cudaError_t e = tl_last_error;
    tl_last_error = cudaSuccess;
    return e;
}
const char* cudaGetErrorString(cudaError_t e) {
    static thread_local char buf[64];
    snprintf(buf, sizeof(buf), "emulated CUDA runtime: error %d", (int)e);
    return buf;
}

// ---- memory -------------------------------------------------------------------------------------------
cudaError_t cudaMalloc(void** p, size_t bytes) {
    *p = nullptr;
    // failure injection for the tests: PHJ_EMU_FAIL_MALLOC = "<device>:<bytes>" makes allocations of at least that many
    // bytes fail on that device (read on every call, so a test can switch it on and off)
    if (const char* f = getenv("PHJ_EMU_FAIL_MALLOC")) {
        int dev = -1;
        unsigned long long min_bytes = 0;
        if (sscanf(f, "%d:%llu", &dev, &min_bytes) == 2 && dev == tl_device && bytes >= min_bytes)
            return cudaErrorMemoryAllocation;
    }
    if (posix_memalign(p, 256, bytes ? bytes : 1) != 0) return cudaErrorMemoryAllocation;
    memset(*p, 0xA5, bytes < 4096 ? bytes : 4096);  // device memory is not zeroed: make the start of it visibly so
    return cudaSuccess;
}
cudaError_t cudaFree(void* p) {
    free(p);
    return cudaSuccess;
}
cudaError_t cudaMallocHost(void** p, size_t bytes) { return cudaMalloc(p, bytes); }
cudaError_t cudaHostAlloc(void** p, size_t bytes, unsigned) { return cudaMalloc(p, bytes); }
cudaError_t cudaFreeHost(void* p) { return cudaFree(p); }
cudaError_t cudaMemcpy(void* dst, const void* src, size_t bytes, cudaMemcpyKind) {
    if (bytes) memmove(dst, src, bytes);
    return cudaSuccess;
}
cudaError_t cudaMemcpyAsync(void* dst, const void* src, size_t bytes, cudaMemcpyKind k, cudaStream_t) {
    return cudaMemcpy(dst, src, bytes, k);
}
cudaError_t cudaMemset(void* p, int v, size_t bytes) {
    if (bytes) memset(p, v, bytes);
    return cudaSuccess;
}
cudaError_t cudaMemsetAsync(void* p, int v, size_t bytes, cudaStream_t) { return cudaMemset(p, v, bytes); }
// One process only: the "IPC handle" of an allocation is its address.
cudaError_t cudaIpcGetMemHandle(cudaIpcMemHandle_t* h, void* p) {
    memset(h, 0, sizeof(*h));
    memcpy(h, &p, sizeof(p));
    return cudaSuccess;
}
cudaError_t cudaIpcOpenMemHandle(void** p, cudaIpcMemHandle_t h, unsigned) {
    memcpy(p, &h, sizeof(*p));
    return cudaSuccess;
}
cudaError_t cudaIpcCloseMemHandle(void*) { return cudaSuccess; }

// ---- streams and events -------------------------------------------------------------------------------
cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, unsigned) {
    *s = (cudaStream_t) new int(0);
    return cudaSuccess;
}
cudaError_t cudaStreamCreateWithPriority(cudaStream_t* s, unsigned f, int) { return cudaStreamCreateWithFlags(s, f); }
cudaError_t cudaStreamDestroy(cudaStream_t s) {
    delete (int*)s;
    return cudaSuccess;
}
cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned) { return cudaSuccess; }
cudaError_t cudaEventCreate(cudaEvent_t* e) {
    *e = (cudaEvent_t) new FakeEvent;
    return cudaSuccess;
}
cudaError_t cudaEventCreateWithFlags(cudaEvent_t* e, unsigned) { return cudaEventCreate(e); }
cudaError_t cudaEventDestroy(cudaEvent_t e) {
    delete (FakeEvent*)e;
    return cudaSuccess;
}
cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t) {
    ((FakeEvent*)e)->ms = now_ms();
    return cudaSuccess;
}
cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t a, cudaEvent_t b) {
    *ms = (float)(((FakeEvent*)b)->ms - ((FakeEvent*)a)->ms);
    if (*ms < 0) *ms = 0;
    return cudaSuccess;
}

}  // extern "C"
