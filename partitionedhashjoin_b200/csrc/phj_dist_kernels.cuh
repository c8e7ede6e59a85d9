// phj_dist_kernels.cuh -- the two kernels of the sharded join (csrc/phj_dist.inl) that are not part of the
// single-GPU path: dist_layout (where every digit run lands in its owner's window) and dist_pull (copies of
// heavy-hitter build partitions). The others -- radix_histogram_lanes, chunk_scan, radix_scatter, gt_clear,
// pt_build, pt_probe -- live in phj_kernels.cuh.
#pragma once
#include "phj_kernels.cuh"

namespace phj {

// ---- the device-side layout ------------------------------------------------------------------------
// all_sizes[src][rel][digit][K + 1] (every rank's chunk_scan totals, all-gathered): entry c = that rank's tuples of
// (digit, chunk c); the build relation counts as chunk 0. Thread d owns split digit d = owner * d_local + local
// partition. The kernel lays out pieces [which_first, which_last] (0: R; 1 + c: chunk c of S): all of them at once
// when the shard was counted up front, one per launch when the counts arrive piece by piece.
//
// Heavy hitters (default; PHJ_FLAG_NO_HOT_DIGITS turns it off; SURVEY.md 8e "skew caveat"): the sizing pass marks the split digits whose
// probe side alone outweighs a quarter of one rank's fair share (key 1 at Zipf 1.25 is 22 % of S) as HOT. Their
// probe tuples do not travel -- every rank keeps its own in an extra partition behind the partitions it owns --
// and their (small, never skewed) build partition is copied from the owner's window to every other rank after
// R has landed (dist_pull). A rank's window: its d_local owned partitions (a hot one holds only the rank's own
// probe tuples), then one partition per hot digit of another owner, in digit order.
constexpr int kMaxHot = 32;

struct PullDesc {
    const ulonglong2* src;  // a hot digit's build partition in its owner's window, as mapped here
    ulonglong2* dst;        // its place in this rank's build window
    unsigned long long n;
};

struct LayoutParams {
    const uint64_t* all_sizes;
    uint32_t world, rank, ndig, d_local, K;
    uint32_t which_first, which_last;
    ulonglong2* const* peer_build;  // [world]: base of every rank's build window, as mapped HERE
    ulonglong2* const* peer_probe;
    ulonglong2** outd;    // out [(1 + K)][ndig]: destination base of digit d for the R launch (0) and the
                          // S launch of chunk c (1 + c): base + the scatter's (piece-local) cursor is the slot
    uint64_t* lb_build;   // out [np + 1]: boundaries of this rank's partitions in its build window
    uint64_t* lb_probe;   // out [K][np + 1]: ... of chunk c's region of its probe window (absolute)
    uint64_t cap_build, cap_probe;  // this rank's windows, tuples
    uint32_t max_keys;              // largest build partition the tables accept
    unsigned long long* flags;      // [1] += 1 if a window is too small, [2] += oversize partitions
    uint32_t np;                    // partitions of this rank: d_local + the hot digits of other owners
    uint32_t n_hot;
    uint32_t hot[kMaxHot];          // ascending
    PullDesc* pulls;                // out [np - d_local]
};

__global__ void __launch_bounds__(256) dist_layout(LayoutParams p) {
    __shared__ uint64_t tot[256];      // tuples of digit d in this piece that land in its OWNER's window
    __shared__ uint64_t earlier[256];  // probe: the same over the chunks before this one
    __shared__ uint64_t own[256];      // probe: this rank's own tuples of digit d in this piece
    __shared__ uint32_t is_hot[256];
    const uint32_t d = threadIdx.x, K = p.K, stride = K + 1, dl = p.d_local;
    const uint32_t owner = d / dl, first = owner * dl, l = d - first, my_first = p.rank * dl;
    const bool mine = owner == p.rank, last = l + 1 == dl;
    is_hot[d] = 0;
    __syncthreads();
    if (d < p.n_hot) is_hot[p.hot[d]] = 1;
    __syncthreads();
    // this rank's extra partitions: the hot digits of other owners, in digit order
    uint32_t j = 0, n_foreign = 0;
    for (uint32_t i = 0; i < p.n_hot; ++i) {
        const uint32_t e = p.hot[i];
        if (e / dl == p.rank) continue;
        ++n_foreign;
        if (e < d) ++j;
    }
    const bool foreign_hot = d < p.ndig && is_hot[d] && !mine;           // extra partition j of this rank
    const bool final_part = n_foreign ? (foreign_hot && j + 1 == n_foreign) : (mine && last);
    for (uint32_t which = p.which_first; which <= p.which_last; ++which) {
        const bool probe = which > 0;
        const uint32_t c = probe ? which - 1 : 0;
        uint64_t before = 0;  // ... of the source ranks before this one, in the owner's window
        if (which != p.which_first) __syncthreads();
        if (d < p.ndig) {
            const bool stays = probe && is_hot[d];  // a hot digit's probe tuples stay where they are
            uint64_t t = 0, e = 0;
            for (uint32_t src = 0; src < p.world; ++src) {
                if (stays && src != owner) continue;
                const uint64_t* sz = p.all_sizes + ((uint64_t)(src * 2 + (probe ? 1 : 0)) * p.ndig + d) * stride;
                t += sz[c];
                if (src < p.rank && !stays) before += sz[c];
                for (uint32_t cc = 0; cc < c; ++cc) e += sz[cc];
            }
            tot[d] = t;
            earlier[d] = e;
            own[d] = p.all_sizes[((uint64_t)(p.rank * 2 + (probe ? 1 : 0)) * p.ndig + d) * stride + c];
        }
        __syncthreads();
        if (d >= p.ndig) continue;
        if (!probe) {
            uint64_t base = 0;
            for (uint32_t e = first; e < d; ++e) base += tot[e];
            p.outd[d] = p.peer_build[owner] + base + before;
            if (mine) {
                p.lb_build[l] = base;
                if (last) p.lb_build[dl] = base + tot[d];
            }
            uint64_t end = base + tot[d];
            if (foreign_hot) {
                uint64_t eb = 0;  // behind the owned partitions and the extra partitions before this one
                for (uint32_t e = my_first; e < my_first + dl; ++e) eb += tot[e];
                for (uint32_t i = 0; i < p.n_hot; ++i)
                    if (p.hot[i] < d && p.hot[i] / dl != p.rank) eb += tot[p.hot[i]];
                end = eb + tot[d];
                p.lb_build[dl + j + 1] = end;
                PullDesc pd;
                pd.src = p.peer_build[owner] + base;
                pd.dst = p.peer_build[p.rank] + eb;
                pd.n = tot[d];
                p.pulls[j] = pd;
            }
            if ((mine || foreign_hot) && tot[d] > p.max_keys) atomicAdd(&p.flags[2], 1ull);
            if (final_part && end > p.cap_build) atomicAdd(&p.flags[1], 1ull);
            continue;
        }
        // The window this digit's tuples go to: the owner's, or -- hot digit -- this rank's own. Chunk c's region of
        // rank w's probe window starts behind the regions of the chunks before it: w's owned partitions plus w's own
        // tuples of the hot digits it does not own.
        const bool stays = is_hot[d] != 0;
        const uint32_t w = stays ? p.rank : owner, wfirst = w * dl;
        uint64_t region = 0, owned = 0;
        for (uint32_t e = wfirst; e < wfirst + dl; ++e) {
            region += earlier[e];
            owned += tot[e];
        }
        for (uint32_t i = 0; i < p.n_hot; ++i) {
            const uint32_t h = p.hot[i];
            if (h / dl == w) continue;
            const uint64_t* sz = p.all_sizes + ((uint64_t)(w * 2 + 1) * p.ndig + h) * stride;
            for (uint32_t cc = 0; cc < c; ++cc) region += sz[cc];
        }
        uint64_t base, end;
        if (!foreign_hot) {  // among w's owned partitions
            uint64_t pre = 0;
            for (uint32_t e = wfirst; e < d; ++e) pre += tot[e];
            base = region + pre;
            end = base + tot[d];
            if (mine) {
                p.lb_probe[(uint64_t)c * (p.np + 1) + l] = base;
                if (last) p.lb_probe[(uint64_t)c * (p.np + 1) + dl] = end;
            }
        } else {             // extra partition j of this rank: its own tuples of a hot digit
            base = region + owned;
            for (uint32_t i = 0; i < p.n_hot; ++i)
                if (p.hot[i] < d && p.hot[i] / dl != p.rank) base += own[p.hot[i]];
            end = base + own[d];
            p.lb_probe[(uint64_t)c * (p.np + 1) + dl + j + 1] = end;
        }
        p.outd[(uint64_t)(1 + c) * p.ndig + d] = p.peer_probe[w] + base + before;
        if (c + 1 == K && final_part && end > p.cap_probe) atomicAdd(&p.flags[1], 1ull);
    }
}

// Copies the build partitions of the hot digits of other owners into this rank's window (after R's barrier).
__global__ void __launch_bounds__(256) dist_pull(const PullDesc* __restrict__ pulls) {
    const PullDesc pd = pulls[blockIdx.y];
    for (uint64_t i = (uint64_t)blockIdx.x * 256 + threadIdx.x; i < pd.n; i += (uint64_t)gridDim.x * 256)
        pd.dst[i] = pd.src[i];
}

}  // namespace phj
