// emu_core.cpp -- TEST INFRASTRUCTURE: the fiber scheduler behind cuda_emu.h (see the header for the model).
// One OS thread executes one kernel launch at a time: CTA after CTA, the CTA's threads as fibers that run until
// they reach a warp or CTA rendezvous.
#include <dlfcn.h>
#include <execinfo.h>
#include <ucontext.h>
#include <signal.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>
#include <sys/mman.h>

#include <algorithm>
#include <vector>

#include "cuda_emu.h"

namespace emu {

thread_local Ctx ctx;

namespace {

constexpr size_t kStackBytes = 96 * 1024;
constexpr int kMaxThreads = 1024;

enum State : int { kReady, kWaitWarp, kWaitCta, kDone };

struct BulkOp {
    void* dst;
    const void* src;
    uint32_t bytes;
};

struct Fiber {
    void* sp;
    State state;
    uint32_t wait_gen;
    uint3 tid;
    unsigned lane, warp;
    std::vector<BulkOp> bulk;  // cp.async.bulk stores issued and not yet waited for
};

struct Warp {
    uint64_t vals[2][32];
    uint32_t live_at[2];  // lanes that took part in generation g & 1
    uint32_t arrived, live, gen;
};

struct Machine {
    std::vector<Fiber> fibers;
    Warp warps[kMaxThreads / 32];
    char* stacks = nullptr;
    void* sched_sp = nullptr;
    Fiber* cur = nullptr;
    int nthreads = 0, ndone = 0;
    uint32_t cta_arrived = 0, cta_gen = 0;
    void (*tramp)(void**) = nullptr;
    void** args = nullptr;
};

thread_local Machine* tl_machine = nullptr;

extern "C" void emu_switch(void** save_sp, void* load_sp);
asm(R"(
    .text
    .globl emu_switch
    .type emu_switch, @function
emu_switch:
    pushq %rbp
    pushq %rbx
    pushq %r12
    pushq %r13
    pushq %r14
    pushq %r15
    movq %rsp, (%rdi)
    movq %rsi, %rsp
    popq %r15
    popq %r14
    popq %r13
    popq %r12
    popq %rbx
    popq %rbp
    ret
    .size emu_switch, .-emu_switch
)");

void yield_to_scheduler() {
    Machine* m = tl_machine;
    Fiber* f = m->cur;
    emu_switch(&f->sp, m->sched_sp);
    ctx.tid = f->tid;  // back on this fiber
    ctx.lane = f->lane;
}

void warp_release_if_complete(Machine* m, Warp& w) {
    if (w.live && (w.arrived & w.live) == w.live) {
        w.live_at[w.gen & 1] = w.live;
        w.arrived = 0;
        ++w.gen;
    }
}

void cta_release_if_complete(Machine* m) {
    if (m->cta_arrived && m->cta_arrived + (uint32_t)m->ndone >= (uint32_t)m->nthreads) {
        m->cta_arrived = 0;
        ++m->cta_gen;
    }
}

void fiber_main() {
    Machine* m = tl_machine;
    Fiber* f = m->cur;
    ctx.tid = f->tid;
    ctx.lane = f->lane;
    m->tramp(m->args);
    if (!f->bulk.empty()) {
        fprintf(stderr, "emu: thread (%u,%u,%u) of block (%u,%u,%u) exited with %zu bulk stores it never waited for\n",
                f->tid.x, f->tid.y, f->tid.z, ctx.bid.x, ctx.bid.y, ctx.bid.z, f->bulk.size());
        abort();
    }
    // the thread has exited: it counts as arrived at every later rendezvous
    f->state = kDone;
    ++m->ndone;
    Warp& w = m->warps[f->warp];
    w.live &= ~(1u << f->lane);
    w.arrived &= ~(1u << f->lane);
    warp_release_if_complete(m, w);
    cta_release_if_complete(m);
    emu_switch(&f->sp, m->sched_sp);
    fprintf(stderr, "emu: a finished fiber was resumed\n");
    abort();
}

// A crash inside an emulated kernel: say where (CTA, thread) before dying.
void on_segv(int sig, siginfo_t* info, void* uc) {
    Machine* m = tl_machine;
    void* pc = (void*)((ucontext_t*)uc)->uc_mcontext.gregs[REG_RIP];
    Dl_info dl{};
    dladdr(pc, &dl);
    fprintf(stderr, "emu: signal %d at address %p, pc %s+0x%zx", sig, info->si_addr, dl.dli_fname ? dl.dli_fname : "?",
            (size_t)((char*)pc - (char*)dl.dli_fbase));
    if (m && m->cur)
        fprintf(stderr, " in an emulated kernel: block (%u,%u,%u) thread (%u,%u,%u) of %d; its stack is [%p, %p)",
                ctx.bid.x, ctx.bid.y, ctx.bid.z, m->cur->tid.x, m->cur->tid.y, m->cur->tid.z, m->nthreads,
                (void*)(m->stacks + (size_t)(m->cur - m->fibers.data()) * kStackBytes),
                (void*)(m->stacks + (size_t)(m->cur - m->fibers.data() + 1) * kStackBytes));
    fprintf(stderr, "\n");
    void* frames[32];
    backtrace_symbols_fd(frames, backtrace(frames, 32), 2);
    _exit(139);
}

void install_segv_handler() {
    static bool done = false;
    if (done) return;
    done = true;
    static char altstack[64 * 1024];
    stack_t ss{};
    ss.ss_sp = altstack;
    ss.ss_size = sizeof(altstack);
    sigaltstack(&ss, nullptr);
    struct sigaction sa{};
    sa.sa_sigaction = on_segv;
    sa.sa_flags = SA_SIGINFO | SA_ONSTACK;
    sigaction(SIGSEGV, &sa, nullptr);
    sigaction(SIGBUS, &sa, nullptr);
}

}  // namespace

const uint64_t* warp_gather(uint64_t v, uint32_t* live) {
    Machine* m = tl_machine;
    Fiber* f = m->cur;
    Warp& w = m->warps[f->warp];
    const uint32_t g = w.gen;
    w.vals[g & 1][f->lane] = v;
    w.arrived |= 1u << f->lane;
    warp_release_if_complete(m, w);
    if (w.gen == g) {
        f->state = kWaitWarp;
        f->wait_gen = g;
        yield_to_scheduler();
    }
    *live = w.live_at[g & 1];
    return w.vals[g & 1];
}

void bulk_defer(void* gdst, const void* ssrc, uint32_t bytes) {
    tl_machine->cur->bulk.push_back(BulkOp{gdst, ssrc, bytes});
}

void bulk_complete() {
    Fiber* f = tl_machine->cur;
    for (const BulkOp& op : f->bulk) memcpy(op.dst, op.src, op.bytes);
    f->bulk.clear();
}

void cta_barrier() {
    Machine* m = tl_machine;
    Fiber* f = m->cur;
    const uint32_t g = m->cta_gen;
    ++m->cta_arrived;
    cta_release_if_complete(m);
    if (m->cta_gen == g) {
        f->state = kWaitCta;
        f->wait_gen = g;
        yield_to_scheduler();
    }
}

// Runs one launch on the calling OS thread. `tramp(args)` calls the kernel with its by-value parameters.
extern "C" __attribute__((visibility("default"))) int emu_run_kernel(void (*tramp)(void**), void** args,
                                                                       unsigned gx, unsigned gy, unsigned gz,
                                                                       unsigned bx, unsigned by, unsigned bz) {
    const int nthreads = (int)(bx * by * bz);
    if (nthreads <= 0 || nthreads > kMaxThreads) return 1;
    if (!tl_machine) {
        if (getenv("PHJ_EMU_TRACE")) install_segv_handler();
        tl_machine = new Machine;
        tl_machine->stacks = (char*)mmap(nullptr, kStackBytes * kMaxThreads, PROT_READ | PROT_WRITE,
                                         MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
        if (tl_machine->stacks == (char*)MAP_FAILED) return 2;
        tl_machine->fibers.resize(kMaxThreads);
    }
    Machine* m = tl_machine;
    if (m->cur) {
        fprintf(stderr, "emu: nested kernel launch\n");
        abort();
    }
    m->tramp = tramp;
    m->args = args;
    m->nthreads = nthreads;
    const Ctx saved = ctx;
    ctx.gdim = dim3(gx, gy, gz);
    ctx.bdim = dim3(bx, by, bz);
    for (unsigned z = 0; z < gz; ++z)
        for (unsigned y = 0; y < gy; ++y)
            for (unsigned x = 0; x < gx; ++x) {
                ctx.bid = uint3{x, y, z};
                m->ndone = 0;
                m->cta_arrived = 0;
                const int nwarps = (nthreads + 31) / 32;
                for (int wi = 0; wi < nwarps; ++wi) {
                    Warp& w = m->warps[wi];
                    w.arrived = 0;
                    w.gen = 0;
                    const int lanes = nthreads - wi * 32 >= 32 ? 32 : nthreads - wi * 32;
                    w.live = lanes == 32 ? 0xffffffffu : (1u << lanes) - 1u;
                }
                for (int t = 0; t < nthreads; ++t) {
                    Fiber& f = m->fibers[t];
                    f.state = kReady;
                    f.bulk.clear();
                    f.tid = uint3{(unsigned)t % bx, ((unsigned)t / bx) % by, (unsigned)t / (bx * by)};
                    f.lane = (unsigned)t & 31;
                    f.warp = (unsigned)t >> 5;
                    // initial stack: six callee-saved registers, the entry point, a slot that keeps the ABI alignment
                    uint64_t* top = (uint64_t*)(m->stacks + (size_t)(t + 1) * kStackBytes);
                    top[-1] = 0;
                    top[-2] = (uint64_t)(uintptr_t)&fiber_main;
                    for (int i = 3; i <= 8; ++i) top[-i] = 0;
                    f.sp = top - 8;
                }
                // Order in which runnable fibers get their turn. Any order is a legal execution (threads only meet at
                // rendezvous points), so a kernel that passes only in one of them relies on something CUDA does not
                // promise: PHJ_EMU_SCHED = forward (default) | reverse | random (a new permutation every sweep).
                static const int sched = [] {
                    const char* v = getenv("PHJ_EMU_SCHED");
                    return !v ? 0 : !strcmp(v, "reverse") ? 1 : !strcmp(v, "random") ? 2 : 0;
                }();
                uint64_t rnd = 0x9E3779B97F4A7C15ull * (x + 1) + y * 1315423911ull + z;
                while (m->ndone < nthreads) {
                    bool progress = false;
                    if (sched == 2) {  // a cheap permutation: t -> (a * t + b) mod nthreads with a odd (nthreads is even
                        rnd = rnd * 6364136223846793005ull + 1442695040888963407ull;  // or 1 .. 1024: gcd(a, n) = 1
                    }                                                                  // for powers of two and most n)
                    const uint32_t a = sched == 2 ? (uint32_t)(rnd >> 33) | 1u : 1u, b = sched == 2 ? (uint32_t)(rnd >> 13) : 0u;
                    const bool perm_ok = sched != 2 || std::__gcd<uint32_t>(a % (uint32_t)nthreads, (uint32_t)nthreads) == 1;
                    for (int i = 0; i < nthreads; ++i) {
                        const int t = sched == 1 ? nthreads - 1 - i
                                      : (sched == 2 && perm_ok) ? (int)(((uint64_t)a * (uint32_t)i + b) % (uint32_t)nthreads) : i;
                        Fiber& f = m->fibers[t];
                        if (f.state == kDone) continue;
                        if (f.state == kWaitWarp && m->warps[f.warp].gen == f.wait_gen) continue;
                        if (f.state == kWaitCta && m->cta_gen == f.wait_gen) continue;
                        f.state = kReady;
                        m->cur = &f;
                        emu_switch(&m->sched_sp, f.sp);
                        progress = true;
                    }
                    if (!progress) {
                        fprintf(stderr, "emu: deadlock in CTA (%u,%u,%u): %d of %d threads done, the others wait at a "
                                        "rendezvous that cannot complete\n", x, y, z, m->ndone, nthreads);
                        abort();
                    }
                }
                m->cur = nullptr;
            }
    ctx = saved;
    return 0;
}

}  // namespace emu
