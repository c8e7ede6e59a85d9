"""Multi-GPU radix join: one process per GPU, torch.distributed for the plumbing.

The reference is single-process (no communication backend at all); radix
partitions are independent join units, so the join shards by partition with ONE exchange step:

  1. every rank holds a row-range shard of R and of S (phj_upload);
  2. split   -- PHJ_ALGO_SHARD_SPLIT: histogram / scan / scatter of the shard by owner rank,
                owner = (hash >> SHARD_SHIFT) % world (hash bits disjoint from the local join's);
  3. exchange-- the per-owner pieces travel with one all-to-all per relation (NCCL over NVLink;
                sizes first, a 2*world int64 all-to-all);
  4. local   -- an ordinary PHJ_ALGO_RADIX_PARTITIONING join on what arrived (phj_bind_device on
                the receive buffers, zero-copy);
  5. reduce  -- all-reduce (sum) of the match counts.

Equal keys hash to the same owner, so the sum of the local counts is the global count. The host
logic below is backend-agnostic: `GpuBackend` drives libphj_b200.so; the tests run the same
orchestration over gloo with a CPU stand-in (tests/test_distributed.py).
"""
from __future__ import annotations

import os
import time

import numpy as np

SHARD_SHIFT = 32  # owner digit = hash bits [32, 32 + log2 world): disjoint from the <= 16 partition bits


class _CudaView:
    """Zero-copy torch view of a raw device pointer (via __cuda_array_interface__)."""

    def __init__(self, ptr: int, rows: int):
        self.__cuda_array_interface__ = {"shape": (rows, 2), "typestr": "<i8", "data": (ptr, False), "version": 3,
                                         "strides": None}


class GpuBackend:
    """split / local join on the device through the C ABI (no CPU fallback)."""

    def __init__(self, world, device, partitions_local=0, hash="xxh3", hash_seed=0x9E3779B97F4A7C15):
        import torch

        from . import engine
        self.torch, self.world, self.device = torch, world, device
        self.split_engine = engine.Engine("shard-split", partitions=world, hash=hash, hash_seed=hash_seed,
                                          device=device, shard_shift=SHARD_SHIFT)
        self.local_engine = engine.Engine("radix-partitioning", partitions=partitions_local, hash=hash,
                                          hash_seed=hash_seed, device=device)
        self._recv = [None, None]
        self.launches = 0

    def upload(self, R, S):
        self.split_engine.upload(R, S)

    def bind_device(self, d_build, n_build, d_probe, n_probe, keepalive=None):
        """Shards that already live on this rank's GPU (e.g. DeviceTuples from the device generators)."""
        self.split_engine.bind_device(d_build, n_build, d_probe, n_probe, keepalive=keepalive)

    def split(self):
        """-> ([send tensor R, send tensor S] ordered by owner, counts[2][world], device ns)."""
        res = self.split_engine.join()
        self.launches += res["kernel_launches"]
        sends, counts = [], []
        for which in (0, 1):
            ptr, _, n = self.split_engine.device_partitions(which)
            b = self.split_engine.read_bounds(which, self.world).astype(np.int64)
            counts.append(np.diff(b))
            if n == 0:
                sends.append(self.torch.empty((0, 2), dtype=self.torch.int64, device=f"cuda:{self.device}"))
            else:
                sends.append(self.torch.as_tensor(_CudaView(ptr, n), device=f"cuda:{self.device}"))
        return sends, np.stack(counts), res["total_ns"]

    def recv_buffer(self, which, rows):
        buf = self._recv[which]
        if buf is None or buf.shape[0] < rows:
            buf = self.torch.empty((max(rows, 1) + max(rows, 1) // 16, 2), dtype=self.torch.int64,
                                   device=f"cuda:{self.device}")
            self._recv[which] = buf
        return buf[:rows]

    def local_join(self, recv_R, recv_S):
        self.torch.cuda.current_stream().synchronize()  # the all-to-all has landed
        self.local_engine.bind_device(recv_R.data_ptr() if recv_R.shape[0] else 0, recv_R.shape[0],
                                      recv_S.data_ptr() if recv_S.shape[0] else 0, recv_S.shape[0],
                                      keepalive=(recv_R, recv_S))
        res = self.local_engine.join()
        self.launches += res["kernel_launches"]
        return res["matches"], res

    def count_tensor(self, value):
        return self.torch.tensor([value], dtype=self.torch.int64, device=f"cuda:{self.device}")

    def int_tensor(self, array):
        return self.torch.as_tensor(np.ascontiguousarray(array, dtype=np.int64), device=f"cuda:{self.device}")

    def close(self):
        self.split_engine.close()
        self.local_engine.close()


class ShardedRadixJoin:
    """One rank of the sharded join. `dist` is torch.distributed (already initialised) or None for a
    single rank."""

    def __init__(self, dist, rank, world, backend):
        if world & (world - 1):
            raise ValueError("the number of ranks must be a power of two")
        self.dist, self.rank, self.world, self.backend = dist, rank, world, backend
        self.last = {}

    def upload(self, R_shard, S_shard):
        self.backend.upload(R_shard, S_shard)

    def join(self) -> dict:
        dist, be, world = self.dist, self.backend, self.world
        t0 = time.perf_counter()
        sends, counts, split_ns = be.split()                      # counts[rel][owner]
        t1 = time.perf_counter()
        if world > 1:
            send_counts = be.int_tensor(counts.T.reshape(-1))     # [owner][rel]
            recv_counts = be.int_tensor(np.zeros(2 * world))
            dist.all_to_all_single(recv_counts, send_counts)
            rc = recv_counts.cpu().numpy().reshape(world, 2)      # [source][rel]
            recvs = []
            for which in (0, 1):
                out_rows = rc[:, which].tolist()
                buf = be.recv_buffer(which, int(sum(out_rows)))
                dist.all_to_all_single(buf, sends[which], output_split_sizes=out_rows,
                                       input_split_sizes=counts[which].tolist())
                recvs.append(buf)
        else:
            recvs = sends
        t2 = time.perf_counter()
        local_matches, res = be.local_join(recvs[0], recvs[1])
        t3 = time.perf_counter()
        total = be.count_tensor(local_matches)
        if world > 1:
            dist.all_reduce(total)
        matches = int(total.item())
        t4 = time.perf_counter()
        self.last = {"matches": matches, "local_matches": int(local_matches), "split_s": t1 - t0,
                     "exchange_s": t2 - t1, "local_s": t3 - t2, "reduce_s": t4 - t3, "total_s": t4 - t0,
                     "recv_rows": [int(r.shape[0]) for r in recvs], "send_bytes_remote":
                     int(16 * (counts.sum() - counts[:, self.rank].sum())) if world > 1 else 0,
                     "local_result": res, "split_device_ns": split_ns}
        return self.last

    def close(self):
        self.backend.close()


class NpjGpuBackend:
    """Device side of the multi-GPU no-partitioning join: the rank's shards as device tensors and a
    PHJ_ALGO_NO_PARTITIONING engine that joins the gathered build relation with the local probe shard."""

    def __init__(self, world, device, hash="xxh3", hash_seed=0x9E3779B97F4A7C15, chained_table=False):
        import torch

        from . import _lib, engine
        self.torch, self.world, self.device = torch, world, device
        self.engine = engine.Engine("no-partitioning", hash=hash, hash_seed=hash_seed, device=device,
                                    flags=_lib.FLAG_CHAINED_TABLE if chained_table else 0)
        self._full = None
        self.launches = 0

    def upload(self, R, S):
        from .engine import as_tuples
        dev = f"cuda:{self.device}"
        to_dev = lambda rel: self.torch.from_numpy(as_tuples(rel).view("<i8").reshape(-1, 2)).to(dev)
        self.R, self.S = to_dev(R), to_dev(S)

    def build_shard(self):
        return self.R

    def gather_buffer(self, rows):
        if self._full is None or self._full.shape[0] < rows:
            self._full = self.torch.empty((max(rows, 1), 2), dtype=self.torch.int64, device=f"cuda:{self.device}")
        return self._full[:rows]

    def local_join(self, full_R):
        self.torch.cuda.current_stream().synchronize()  # the broadcasts have landed
        self.engine.bind_device(full_R.data_ptr() if full_R.shape[0] else 0, full_R.shape[0],
                                self.S.data_ptr() if self.S.shape[0] else 0, self.S.shape[0],
                                keepalive=(full_R, self.S))
        res = self.engine.join()
        self.launches += res["kernel_launches"]
        return res["matches"], res

    def count_tensor(self, value):
        return self.torch.tensor([value], dtype=self.torch.int64, device=f"cuda:{self.device}")

    def int_tensor(self, array):
        return self.torch.as_tensor(np.ascontiguousarray(array, dtype=np.int64), device=f"cuda:{self.device}")

    def close(self):
        self.engine.close()


class ReplicatedNoPartitioningJoin:
    """The no-partitioning join on several GPUs (last row): no shuffle. Every rank holds a
    row shard of R and of S; the build shards are gathered so that every rank builds the table of the
    WHOLE build relation (NoPartitioning::HashJoiner::Build, src/NoPartitioning/HashJoin.hpp:76-126,
    replicated), probes it with its own shard of S (Probe, :128-187) and the counts are summed. Probe
    tuples never move; the traffic is |R| x (world - 1) tuples per rank, once per join."""

    def __init__(self, dist, rank, world, backend):
        self.dist, self.rank, self.world, self.backend = dist, rank, world, backend
        self.last = {}

    def upload(self, R_shard, S_shard):
        self.backend.upload(R_shard, S_shard)

    def join(self) -> dict:
        dist, be, world = self.dist, self.backend, self.world
        t0 = time.perf_counter()
        shard = be.build_shard()
        if world > 1:
            sizes = [be.int_tensor(np.zeros(1)) for _ in range(world)]
            dist.all_gather(sizes, be.int_tensor(np.array([shard.shape[0]])))
            rows = [int(x.item()) for x in sizes]
            full = be.gather_buffer(sum(rows))
            first = 0
            for src, n in enumerate(rows):  # uneven shards: one broadcast per source into its slice
                piece = full[first:first + n]
                if src == self.rank:
                    piece.copy_(shard)
                if n:
                    dist.broadcast(piece, src=src)
                first += n
        else:
            rows, full = [int(shard.shape[0])], shard
        t1 = time.perf_counter()
        local_matches, res = be.local_join(full)
        t2 = time.perf_counter()
        total = be.count_tensor(local_matches)
        if world > 1:
            dist.all_reduce(total)
        matches = int(total.item())
        t3 = time.perf_counter()
        self.last = {"matches": matches, "local_matches": int(local_matches), "build_rows": int(sum(rows)),
                     "split_s": 0.0, "exchange_s": t1 - t0, "local_s": t2 - t1, "reduce_s": t3 - t2,
                     "total_s": t3 - t0, "recv_bytes": int(16 * (sum(rows) - rows[self.rank])) if world > 1 else 0,
                     "send_bytes_remote": int(16 * rows[self.rank] * (world - 1)), "local_result": res}
        return self.last

    def close(self):
        self.backend.close()


def split_plan(world, partitions_local, pass1_in_shuffle=True):
    """(b1, b2) of the local radix join and the number of split digits. With pass1_in_shuffle the
    split digit is owner rank x local pass-1 digit (<= 256 digits in all), so what a rank receives
    is already pass-1 partitioned; otherwise the split is by owner rank only (b1 = b2 = 0 here)."""
    bits = int(partitions_local).bit_length() - 1
    if not pass1_in_shuffle or partitions_local < 4 or (1 << bits) != partitions_local:
        return 0, 0, world
    b1 = min((bits + 1) // 2, 8 - (world.bit_length() - 1))
    if os.environ.get("PHJ_SPLIT_B1"):  # experiments: fewer split digits = longer NVLink runs, wider pass 2
        b1 = min(b1, int(os.environ["PHJ_SPLIT_B1"]))
    b2 = bits - b1
    if b1 < 1 or b2 > 8:
        return 0, 0, world
    return b1, b2, world << b1


class FusedGpuBackend(GpuBackend):
    """The partition shuffle as NVLink stores: every rank's receive window is mapped by its peers
    (CUDA IPC, phj_shared_*), and the split scatter's TMA bulk stores write each piece straight
    into its owner's window (phj_shard_scatter). No separate all-to-all of tuples, and -- with
    pass1_in_shuffle -- no local pass 1 either: the split digit is (owner rank : local pass-1
    digit), the window is laid out digit-major, and the local join starts at pass 2."""

    def __init__(self, world, device, partitions_local=4096, hash="xxh3", hash_seed=0x9E3779B97F4A7C15,
                 pass1_in_shuffle=True):
        import ctypes
        import torch

        from . import _lib, engine
        self.torch, self.world, self.device = torch, world, device
        self.b1, self.b2, self.ndig = split_plan(world, partitions_local, pass1_in_shuffle)
        self.d1 = self.ndig // world
        if self.b1:
            self.split_engine = engine.Engine("shard-split", partitions=self.ndig, hash=hash, hash_seed=hash_seed,
                                              device=device, shard_shift=self.b2,
                                              flags=int(os.environ.get("PHJ_SPLIT_FLAGS", "0"), 0))
            self.local_engine = engine.Engine("radix-partitioning", partitions=partitions_local,
                                              radix_bits=(self.b1, self.b2), hash=hash, hash_seed=hash_seed,
                                              device=device)
        else:
            self.split_engine = engine.Engine("shard-split", partitions=world, hash=hash, hash_seed=hash_seed,
                                              device=device, shard_shift=SHARD_SHIFT)
            self.local_engine = engine.Engine("radix-partitioning", partitions=partitions_local, hash=hash,
                                              hash_seed=hash_seed, device=device)
        self._recv = [None, None]
        self.launches = 0
        self._C, self._lib, self._check = ctypes, _lib.lib, _lib.check
        self.win = [None] * 3                                  # own windows: device pointers
        self.peer = [[None] * world for _ in range(3)]         # mapped windows of every rank

    # -- windows ---------------------------------------------------------------------------
    def win_alloc(self, which, rows):
        C = self._C
        ptr, handle = C.c_void_p(), (C.c_ubyte * 64)()
        self._check(self._lib.phj_shared_alloc(self.device, max(rows, 1) * 16, C.byref(ptr), handle))
        self.win[which] = ptr.value
        return bytes(handle)

    def win_free(self, which):
        if self.win[which]:
            self._check(self._lib.phj_shared_free(self.device, self._C.c_void_p(self.win[which])))
            self.win[which] = None

    def peer_open(self, which, src, handle, own):
        C = self._C
        if own:
            self.peer[which][src] = self.win[which]
            return
        ptr = C.c_void_p()
        buf = (C.c_ubyte * 64).from_buffer_copy(handle)
        self._check(self._lib.phj_shared_open(self.device, buf, C.byref(ptr)))
        self.peer[which][src] = ptr.value

    def peer_close(self, which, src, own):
        p = self.peer[which][src]
        if p and not own:
            self._check(self._lib.phj_shared_close(self.device, self._C.c_void_p(p)))
        self.peer[which][src] = None

    # -- the split, in two halves ------------------------------------------------------------
    def count(self):
        """Tuples of this rank's shard per split digit: counts[rel][digit]."""
        counts = np.zeros((2, self.ndig), dtype=np.uint64)
        self._check(self._lib.phj_shard_count(self.split_engine._h, counts.ctypes.data))
        return counts.astype(np.int64)

    def scatter(self, offsets, owner_of):
        """offsets[rel][digit]: first row of this rank's piece of `digit` inside the window of
        owner_of[digit] (or owner_of[rel][digit] when the two relations go different ways)."""
        from ._lib import PhjResult
        C = self._C
        owner_of = np.asarray(owner_of)
        if owner_of.ndim == 1:
            owner_of = np.stack([owner_of, owner_of])
        arrs = []
        for which in (0, 1):
            ptrs = (C.c_void_p * self.ndig)(*[C.c_void_p(self.peer[which][int(owner_of[which][d])])
                                              for d in range(self.ndig)])
            offs = np.ascontiguousarray(offsets[which], dtype=np.uint64)
            arrs += [ptrs, offs]
        res = PhjResult()
        self._check(self._lib.phj_shard_scatter(self.split_engine._h, 0, arrs[0], arrs[1].ctypes.data, arrs[2],
                                                arrs[3].ctypes.data, C.byref(res)))
        self.launches += res.kernel_launches
        return int(res.total_ns)

    def pull(self, which, owner, src_row, dst_row, rows):
        """Copy `rows` tuples from rank `owner`'s window (mapped here) into this rank's own window."""
        if rows:
            self._check(self._lib.phj_memcpy_d2d(self.device, self._C.c_void_p(self.win[which] + 16 * dst_row),
                                                 self._C.c_void_p(self.peer[which][owner] + 16 * src_row), 16 * rows))

    def local_join_window(self, rows, bounds, first_digit=0, parent_digits=None):
        """rows[rel] tuples have landed in this rank's windows; bounds[rel] are the boundaries of the
        split digits first_digit, first_digit + 1, ... this rank owns, in window order (or of the
        explicitly listed parent_digits)."""
        ptr = [self.win[w] if rows[w] else 0 for w in (0, 1)]
        if self.b1 and len(bounds[0]) < 2:
            return 0, {"kernel_launches": 0, "hbm_bytes_alg": 0, "total_ns": 1, "d2h_bytes": 0}  # owns no digit
        if self.b1:
            self.local_engine.bind_device_partitioned(ptr[0], rows[0], ptr[1], rows[1], bounds[0], bounds[1],
                                                      first_parent=first_digit, parent_space=self.ndig,
                                                      parent_digits=parent_digits)
        else:
            self.local_engine.bind_device(ptr[0], rows[0], ptr[1], rows[1])
        res = self.local_engine.join()
        self.launches += res["kernel_launches"]
        return res["matches"], res


class FusedShardedRadixJoin:
    """One rank of the sharded join with the shuffle fused into the split scatter.

      count     histogram + scan of the shard by split digit = (owner rank : local pass-1 digit)
                                                                          (phj_shard_count)
      sizes     all-gather of the world x 2 x digits piece sizes: every rank knows where each of its
                pieces starts inside its owner's window (digit-major, then source rank: the stable
                order of the reference's partitionTable), and how big every window must be
      windows   (re)allocated and re-mapped collectively only when a window is too small
      scatter   the split scatter writes every piece into its owner's window over NVLink
      barrier   all pieces have landed
      local     radix join on the window, starting at pass 2; all-reduce of the count
    """

    GROW = 1.125

    def __init__(self, dist, rank, world, backend, balance=True):
        if world & (world - 1):
            raise ValueError("the number of ranks must be a power of two")
        self.dist, self.rank, self.world, self.backend, self.balance = dist, rank, world, backend, balance
        if backend is not None:
            backend.rank = rank
        self.caps = np.zeros((3, world), dtype=np.int64)  # rows of every rank's windows (same on all ranks)
        self.last = {}

    def upload(self, R_shard, S_shard):
        self.backend.upload(R_shard, S_shard)

    def _gather_counts(self, counts):
        if self.world == 1:
            return counts[None]
        be = self.backend
        mine = be.int_tensor(counts.reshape(-1))
        everyone = be.int_tensor(np.zeros(self.world * counts.size))
        self.dist.all_gather_into_tensor(everyone, mine)
        return everyone.cpu().numpy().reshape(self.world, 2, -1)  # [source][rel][digit]

    def _ensure_windows(self, need):
        """need: {window id: rows arriving at every owner}. Every rank sees the same `need` and `caps`,
        so all of them take the same decision without further communication."""
        grow = {w: np.asarray(n) > self.caps[w] for w, n in need.items()}
        if not any(g.any() for g in grow.values()):
            return False
        be, world, rank = self.backend, self.world, self.rank
        for w, g in grow.items():
            for r in range(world):
                if g[r]:
                    be.peer_close(w, r, own=(r == rank))
        if world > 1:
            self.dist.barrier()  # nobody maps a window that is about to be freed
        handles = {}
        for w, g in grow.items():
            self.caps[w] = np.where(g, (np.asarray(need[w]) * self.GROW).astype(np.int64) + 4096, self.caps[w])
            if g[rank]:
                be.win_free(w)
                handles[w] = be.win_alloc(w, int(self.caps[w][rank]))
        if world > 1:
            everyone = [None] * world
            self.dist.all_gather_object(everyone, handles)
        else:
            everyone = [handles]
        for w, g in grow.items():
            for r in range(world):
                if g[r]:
                    be.peer_open(w, r, everyone[r][w], own=(r == rank))
        return True

    @staticmethod
    def ownership(M, world, balance=True):
        """first[o] = first split digit owned by rank o (first[world] = digits). Equal digit ranges, or
        -- balance -- contiguous ranges cut where the cumulative tuple count crosses o / world of the
        total, so that a rank whose range holds heavy-hitter digits owns fewer of them (SURVEY 8e
        "skew caveat"). M[source][rel][digit]; every rank computes the same cut from the same M."""
        ndig = M.shape[2]
        if not balance or world == 1 or ndig == world:
            return np.arange(world + 1, dtype=np.int64) * (ndig // world)
        cum = np.cumsum(M.sum(axis=(0, 1)))
        first = [0]
        for o in range(1, world):
            target = cum[-1] * o / world
            i = int(np.searchsorted(cum, target, side="left"))          # cum[i] >= target
            below = cum[i - 1] if i > 0 else 0
            b = i + 1 if i < ndig and abs(cum[i] - target) < abs(target - below) else i
            first.append(min(max(b, first[-1]), ndig))
        first.append(ndig)
        return np.asarray(first, dtype=np.int64)

    @staticmethod
    def hot_digits(M, world):
        """Split digits whose probe side alone is heavier than a quarter of one rank's fair share of the
        probe relation (never the case for uniform keys: a digit is at most 1/32 of a share): an
        ownership cut cannot balance those (key 1 at Zipf 1.25 is 22 % of S). Their probe tuples stay
        on the rank that holds them -- no NVLink traffic at all -- and the (small, never skewed) build
        side of the digit is replicated to every rank instead: SURVEY 8e "skew caveat"."""
        tot_s = M[:, 1, :].sum(axis=0)
        if world == 1 or M.shape[2] == world:
            return np.zeros(M.shape[2], dtype=bool)
        return tot_s * world * 4 > max(int(tot_s.sum()), 1)   # heavier than a quarter of a fair share

    @staticmethod
    def layout_hot(M, world, rank, first, hot):
        """layout() with replicated heavy-hitter digits. Window of rank o: its owned digits first
        (build side from every source; probe side from every source, but only o's own piece of a hot
        digit), then one slot per hot digit owned by another rank (build side pulled from the owner
        after the exchange, probe side = o's own piece). Returns rows per window, this rank's write
        offsets and destination ranks per relation, its local boundaries, the digit of every local
        parent, and the pulls (owner, source row, destination row, rows) of the build side."""
        ndig = M.shape[2]
        owner = np.searchsorted(first[1:], np.arange(ndig), side="right")
        tot = M.sum(axis=0)
        hot_list = [int(d) for d in np.nonzero(hot)[0]]
        need = np.zeros((2, world), dtype=np.int64)
        base = np.full((2, world, ndig), -1, dtype=np.int64)
        parents = []
        for o in range(world):
            plist = list(range(int(first[o]), int(first[o + 1]))) + [d for d in hot_list if owner[d] != o]
            pos = [0, 0]
            for d in plist:
                base[0, o, d], base[1, o, d] = pos
                pos[0] += int(tot[0, d])
                pos[1] += int(M[o, 1, d]) if hot[d] else int(tot[1, d])
            need[:, o] = pos
            parents.append(plist)
        before = M[:rank].sum(axis=0)                          # lower ranks' pieces come first
        dst = np.stack([owner, np.where(hot, rank, owner)])    # hot probe pieces stay here
        offsets = np.zeros((2, ndig), dtype=np.int64)
        for d in range(ndig):
            offsets[0, d] = base[0, owner[d], d] + before[0, d]
            offsets[1, d] = base[1, rank, d] if hot[d] else base[1, owner[d], d] + before[1, d]
        mine = parents[rank]
        bounds = np.array([[base[rel, rank, d] for d in mine] + [need[rel, rank]] for rel in (0, 1)], dtype=np.int64)
        pulls = [(int(owner[d]), int(base[0, owner[d], d]), int(base[0, rank, d]), int(tot[0, d]))
                 for d in mine if owner[d] != rank]
        return need, offsets, dst, bounds, np.asarray(mine, dtype=np.int64), pulls

    @staticmethod
    def layout(M, world, rank, first=None):
        """From M[source][rel][digit] and the ownership cut: rows arriving at every owner, this rank's
        write offsets per digit, the boundaries of this rank's own digits inside its windows, and the
        owner of every digit."""
        ndig = M.shape[2]
        if first is None:
            first = np.arange(world + 1, dtype=np.int64) * (ndig // world)
        owner_of = np.searchsorted(first[1:], np.arange(ndig), side="right")
        tot = M.sum(axis=0)                                     # [rel][digit]
        excl = np.cumsum(tot, axis=1) - tot                     # global exclusive prefix over digits
        start = np.concatenate([excl, tot.sum(axis=1, keepdims=True)], axis=1)[:, first]  # [rel][world + 1]
        need = np.diff(start, axis=1)                           # [rel][owner]
        base = excl - start[:, owner_of]                        # start of a digit inside its owner's window
        offsets = base + M[:rank].sum(axis=0)                   # lower ranks' pieces come first
        lo, hi = int(first[rank]), int(first[rank + 1])
        bounds = np.concatenate([base[:, lo:hi], need[:, rank:rank + 1]], axis=1) if hi > lo else \
            np.zeros((2, 1), dtype=np.int64)
        return need, offsets, bounds, owner_of

    def join(self) -> dict:
        be, world, rank = self.backend, self.world, self.rank
        t0 = time.perf_counter()
        counts = be.count()                                   # [rel][digit]
        t1 = time.perf_counter()
        M = self._gather_counts(counts)                       # [source][rel][digit]
        hot = self.hot_digits(M, world) if self.balance and getattr(be, "b1", 0) else np.zeros(M.shape[2], dtype=bool)
        if hot.any():
            # heavy hitters: their probe tuples do not travel, their build side is replicated
            W = M.copy()
            W[:, 1, hot] = 0
            first = self.ownership(W, world, True)
            need, offsets, owner_of, bounds, parent_digits, pulls = self.layout_hot(M, world, rank, first, hot)
        else:
            first = self.ownership(M, world, self.balance)
            need, offsets, bounds, owner_of = self.layout(M, world, rank, first)
            parent_digits, pulls = None, []
        regrown = self._ensure_windows({0: need[0], 1: need[1]})
        t2 = time.perf_counter()
        scatter_ns = be.scatter(offsets, owner_of)
        if world > 1:
            self.dist.barrier()                               # every rank's stores are complete
        for owner, src_row, dst_row, nrows in pulls:          # replicate the hot digits' build side
            be.pull(0, owner, src_row, dst_row, nrows)
        t3 = time.perf_counter()
        rows = [int(need[0][rank]), int(need[1][rank])]
        local_matches, res = be.local_join_window(rows, bounds, int(first[rank]), parent_digits)
        t4 = time.perf_counter()
        total = be.count_tensor(local_matches)
        if world > 1:
            self.dist.all_reduce(total)
        matches = int(total.item())
        t5 = time.perf_counter()
        stay = np.asarray(owner_of)
        stay = (stay if stay.ndim == 2 else np.stack([stay, stay])) == rank   # pieces that do not leave this GPU
        mine = (counts * stay).sum()
        self.last = {"matches": matches, "local_matches": int(local_matches), "split_s": t1 - t0,
                     "first_digit": first.tolist(), "hot_digits": [int(d) for d in np.nonzero(hot)[0]],
                     "parent_digits": None if parent_digits is None else parent_digits.tolist(),
                     "exchange_s": t3 - t1, "sizes_s": t2 - t1, "scatter_s": t3 - t2, "local_s": t4 - t3,
                     "reduce_s": t5 - t4, "total_s": t5 - t0, "recv_rows": rows, "regrown": regrown,
                     "send_bytes_remote": int(16 * (counts.sum() - mine)) if world > 1 else 0,
                     "local_result": res, "split_device_ns": scatter_ns, "scatter_device_ns": scatter_ns,
                     "bounds": bounds}
        return self.last

    def close(self):
        be, world, rank = self.backend, self.world, self.rank
        for w in range(3):
            for r in range(world):
                be.peer_close(w, r, own=(r == rank))
        if world > 1:
            self.dist.barrier()
        for w in range(3):
            be.win_free(w)
        be.close()


class ShardedJoin:
    """One rank of the sharded radix join that lives INSIDE libphj_b200.so (phj_dist_*, csrc/phj_dist.inl):
    count -> NCCL all-gather of the piece sizes -> device-side layout -> the radix scatter stores R and then S,
    chunk by chunk, straight into the owners' windows over NVLink while the local L2-table probe of the previous
    chunk runs -> NCCL all-reduce of the count. The host enqueues the join and waits once; Python only bootstraps
    (the NCCL unique id travels over torch.distributed) and calls phj_dist_join. `dist` = torch.distributed
    (initialised) or None for a single rank."""

    def __init__(self, dist, rank, world, device, partitions=0, chunks=0, hash="xxh3",
                 hash_seed=0x9E3779B97F4A7C15, table_seed=1, flags=0, split_ctas=0):
        import ctypes as C

        import torch  # noqa: F401  (first: the library then binds the NCCL torch has loaded)

        from . import _lib
        self._C, self._lib, self._check = C, _lib.lib, _lib.check
        self.dist, self.rank, self.world, self.device = dist, rank, world, device
        ident = (C.c_ubyte * 128)()
        if rank == 0:
            self._check(self._lib.phj_nccl_unique_id(ident))
        if world > 1:
            box = [bytes(ident)]
            dist.broadcast_object_list(box, src=0)
            ident = (C.c_ubyte * 128).from_buffer_copy(box[0])
        cfg = _lib.PhjConfig()
        cfg.algo = _lib.ALGO_RADIX_PARTITIONING
        cfg.hash = _lib.HASH_NAMES[hash] if isinstance(hash, str) else hash
        cfg.partitions = partitions
        cfg.hash_seed = hash_seed & (2**64 - 1)
        cfg.table_seed = table_seed & (2**64 - 1)
        cfg.device = device
        cfg.flags = flags
        cfg.split_chunks = chunks
        cfg.split_ctas = split_ctas  # SMs of the NVLink scatter while it overlaps the local join (0 = choose)
        self._h = C.c_void_p()
        self._check(self._lib.phj_dist_create(C.byref(cfg), rank, world, ident, C.byref(self._h)))
        self._keep = None
        self.launches = 0
        self.last = {}

    def upload(self, R_shard, S_shard):
        from .engine import as_tuples
        R, S = as_tuples(R_shard), as_tuples(S_shard)
        self._check(self._lib.phj_dist_upload(self._h, R.ctypes.data, R.shape[0], S.ctypes.data, S.shape[0]))

    def bind_device(self, d_build, n_build, d_probe, n_probe, keepalive=None):
        C = self._C
        self._check(self._lib.phj_dist_bind_device(self._h, C.c_void_p(d_build), n_build, C.c_void_p(d_probe), n_probe))
        self._keep = keepalive

    def join(self) -> dict:
        from ._lib import PhjResult
        res = PhjResult()
        self._check(self._lib.phj_dist_join(self._h, self._C.byref(res)))
        self.last = res.as_dict()
        self.launches += self.last["kernel_launches"]
        return self.last

    def info(self) -> dict:
        from ._lib import PhjDistLayout
        lay = PhjDistLayout()
        self._check(self._lib.phj_dist_info(self._h, self._C.byref(lay)))
        return {"world": lay.world, "rank": lay.rank, "digits": lay.digits, "local_partitions": lay.local_partitions,
                "chunks": lay.chunks, "region_buckets": lay.region_buckets,
                "window_tuples": [int(lay.window_tuples[0]), int(lay.window_tuples[1])], "resizes": lay.resizes,
                "sent_remote_bytes": int(lay.sent_remote_bytes), "partitions_here": lay.partitions_here,
                "hot_digits": [int(lay.hot_digits[i]) for i in range(lay.hot_count)]}

    def kernel_timing(self, filter):
        """CUDA events around the kernels of the following joins: None = none, "" = all, else names containing it."""
        self._check(self._lib.phj_dist_kernel_timing(self._h, None if filter is None else filter.encode()))

    def kernel_times(self):
        C = self._C
        names, ns = (C.c_char_p * 64)(), (C.c_uint64 * 64)()
        n = self._lib.phj_dist_kernel_times(self._h, names, ns, 64)
        return [(names[i].decode(), int(ns[i])) for i in range(n)]

    def kernel_trace(self):
        """[(kernel, begin ns, end ns)] of the last join relative to its first event (PHJ_KERNEL_TIMES=1)."""
        C = self._C
        names, b, e = (C.c_char_p * 64)(), (C.c_uint64 * 64)(), (C.c_uint64 * 64)()
        n = self._lib.phj_dist_kernel_trace(self._h, names, b, e, 64)
        return [(names[i].decode(), int(b[i]), int(e[i])) for i in range(n)]

    def measure_peer_copy(self, nbytes=1 << 30, repeats=3):
        """GB/s of a plain peer cudaMemcpyAsync of this rank's window into the next rank's, all ranks at once."""
        ns = self._C.c_uint64()
        self._check(self._lib.phj_dist_measure_peer_copy(self._h, nbytes, repeats, self._C.byref(ns)))
        lay = self.info()
        moved = min(nbytes, 16 * lay["window_tuples"][1])
        return moved / ns.value if ns.value else 0.0

    def read_window(self, which):
        """(tuples in this rank's window, boundaries): build -> [partitions_here + 1]; probe ->
        [chunks][partitions_here + 1], absolute positions (test read-back)."""
        from ._lib import TUPLE_DTYPE
        lay = self.info()
        nb = lay["partitions_here"] + 1
        bounds = np.zeros(nb if which == 0 else lay["chunks"] * nb, dtype=np.uint64)
        self._check(self._lib.phj_dist_read_window(self._h, which, None, 0, bounds.ctypes.data))
        out = np.empty(int(bounds[-1]), dtype=TUPLE_DTYPE)
        self._check(self._lib.phj_dist_read_window(self._h, which, out.ctypes.data, out.shape[0], bounds.ctypes.data))
        return out, (bounds if which == 0 else bounds.reshape(lay["chunks"], nb)).astype(np.int64)

    def close(self):
        if getattr(self, "_h", None):
            self._lib.phj_dist_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def shard_inputs(phj, rank, world, n_build, n_probe, skew, base_seed, batches):
    """Weak-scaling shards: R = keys 1..world*n_build cut by rank, S = Zipf over the whole key range."""
    Rp, Sp = phj.PinnedTuples(n_build), phj.PinnedTuples(n_probe)
    phj.fill_sequential(Rp.array, 1 + rank * n_build)
    Rp.array["payload"] += rank * n_build
    phj.fill_zipf(Sp.array, skew, 1, world * n_build, base_seed + 7919 * rank, batches)
    return Rp, Sp


class SingleRank:
    """Stand-in for torch.distributed when there is one rank (bench.py --gpus 1 --workload scaled)."""

    class ReduceOp:
        SUM, MAX, MIN = "sum", "max", "min"

    @staticmethod
    def all_reduce(t, op=None):
        return t

    @staticmethod
    def all_gather(out, t):
        out[0].copy_(t)

    @staticmethod
    def barrier():
        pass

    @staticmethod
    def destroy_process_group():
        pass


def scaled_inputs(phj, dist, torch, rank, world, local, skew):
    """BASELINE.json configs[4]: 160 M x 3.2 B in total, row-sharded over the ranks (strong scaling), generated ON
    the device (the host generator would need minutes and a 51 GB upload). Probe keys are drawn over twice the
    build key range, so the count is not |S|; the independent count (chunk-wise) brings
    the probe keys back to the host 100 M at a time and lets numpy count those inside [1, |R|]."""
    total_build, total_probe = 160_000_000, 3_200_000_000
    n_build, n_probe = total_build // world, total_probe // world
    dR = phj.DeviceTuples(n_build, local).fill_sequential(1 + rank * n_build)
    dS = phj.DeviceTuples(n_probe, local).fill_zipf(skew, 1, 2 * total_build, 12345 + 100_003 * rank, 1 << 16)
    mine, step = 0, 100_000_000
    for first in range(0, n_probe, step):
        ids = dS.download(first, min(step, n_probe - first))["id"]
        mine += int(((ids >= 1) & (ids <= total_build)).sum())
    want = int(round(_allreduce(dist, torch, local, [float(mine)])[0]))
    return dR, dS, n_build, n_probe, want


def scaled_other_config(phj, dist, torch, rank, world, local, args, joins=3):
    """The scaled configuration as an `other_configs` entry of the default run: a few joins, count checked."""
    dR, dS, n_build, n_probe, want = scaled_inputs(phj, dist, torch, rank, world, local, 1.05)
    job = ShardedJoin(dist if world > 1 else None, rank, world, local, hash=args.hash)
    job.bind_device(dR.ptr, n_build, dS.ptr, n_probe, keepalive=(dR, dS))
    got = [job.join()["matches"] for _ in range(2)]
    torch.cuda.synchronize()
    dist.barrier()
    t0 = time.perf_counter()
    for _ in range(joins):
        res = job.join()
    torch.cuda.synchronize()
    dist.barrier()
    dt = _allreduce(dist, torch, local, [time.perf_counter() - t0], dist.ReduceOp.MAX)[0] / joins
    lay = job.info()
    job.close()
    dR.close()
    dS.close()
    return {"workload": "160M x 3200M in total over the ranks (strong scaling), device-generated, Zipf 1.05 over 2x the "
                        "build key range", "ms_per_join": dt * 1e3, "Gtuples_s": round(3.36e9 / dt / 1e9, 2),
            "matches": res["matches"], "want": want, "ok": all(g == want for g in got) and res["matches"] == want,
            "want_how": "numpy count of the probe keys inside [1, |R|], read back from the device 100 M at a time, "
                        "summed over the ranks", "shuffle_ms": res["shuffle_ns"] / 1e6, "count_ms": res["count_ns"] / 1e6,
            "digits": lay["digits"], "chunks": lay["chunks"]}


def shuffle_scatter_dram_traffic(world):
    """DRAM bytes (read + write) of the NVLink scatter over one rank's whole shard, from the ncu capture of that launch
    (tools/ncu_split.py -> profiles/r02r_ncu_split_summary.md: GPU 0 as rank 0 of 8, the other seven ranks' windows on a
    second GPU). The capture stands for 8 ranks: the share written locally is 1 / world. Returns (bytes or None, source)."""
    import os
    import re
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "r02r_ncu_split_summary.md")
    if world != 8 or not os.path.exists(path):
        return None, None
    scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
    total = 0.0
    for line in open(path):
        m = re.match(r"\* DRAM (read|write): ([0-9.]+) (\w+)", line)
        if m and m.group(3) in scale:
            total += float(m.group(2)) * scale[m.group(3)]
    return (total or None), "profiles/r02r_ncu_split_summary.md"


def _allreduce(dist, torch, local, values, op=None):
    t = torch.tensor(values, dtype=torch.float64, device=f"cuda:{local}")
    dist.all_reduce(t, op=op or dist.ReduceOp.SUM)
    return t.tolist()


def parity_check(phj, dist, torch, rank, world, local, make_job, skew, n_build=1_000_000, n_probe=20_000_000):
    """A join whose answer is NOT |S|: probe keys drawn over twice the build key range, one
    build key duplicated on every rank, INT64_MIN on both sides. The all-reduced count of the product is compared
    with an independent one: every rank gathers ALL build keys (torch all_gather) and counts its own probe keys with
    numpy isin; the counts are summed. Also: what landed in the windows is as many tuples as were sent."""
    int64_min = -(2 ** 63)
    R = np.empty(n_build, dtype=phj.TUPLE_DTYPE)
    S = np.empty(n_probe, dtype=phj.TUPLE_DTYPE)
    phj.fill_sequential(R, 1 + rank * n_build)
    phj.fill_zipf(S, skew, 1, 2 * world * n_build, 4242 + 31 * rank, 16)
    R["id"][-1] = 7                      # the same build key once more on every rank (duplicates collapse)
    R["id"][-2] = int64_min if rank == 0 else R["id"][-2]
    S["id"][::1_000_003] = int64_min     # the reserved table marker as a probe key: matches rank 0's build tuple
    S["payload"] += rank * n_probe
    job = make_job()
    job.upload(R, S)
    got = [job.join()["matches"] for _ in range(2)]
    keys = torch.from_numpy(np.ascontiguousarray(R["id"])).to(f"cuda:{local}")
    everyone = [torch.empty_like(keys) for _ in range(world)]
    dist.all_gather(everyone, keys)
    build_keys = np.unique(torch.cat(everyone).cpu().numpy())
    mine = int(np.isin(S["id"], build_keys).sum())
    want = int(round(_allreduce(dist, torch, local, [float(mine)])[0]))
    out = {"count": got[-1], "want": want, "ok": got[0] == want and got[1] == want, "skew": skew,
           "build_per_rank": n_build, "probe_per_rank": n_probe,
           "how": "probe keys over 2x the build key range, a duplicated build key, INT64_MIN on both sides; "
                  "want = sum over ranks of numpy isin(own probe keys, all-gathered build keys)"}
    if hasattr(job, "read_window"):
        landed = [int(job.read_window(w)[1].reshape(-1)[-1]) for w in (0, 1)]
        tot_b, tot_p = (int(v) for v in _allreduce(dist, torch, local, [float(landed[0]), float(landed[1])]))
        hot = job.info().get("hot_digits", [])
        out["window_tuples"] = tot_b + tot_p
        # every probe tuple lands exactly once; so does every build tuple, unless heavy-hitter digits are replicated
        out["ok"] = out["ok"] and tot_p == world * n_probe and \
            (tot_b > world * n_build if hot else tot_b == world * n_build)
        if hot:
            out["hot_digits"] = hot
    job.close()
    return out


def bench(args, dist, rank, world, local, make_inputs, workload_config, ClockSampler, metric, unit, measured_hbm_peak):
    """bench.py --gpus N (N > 1). Default workload: every rank brings a 10 M x 200 M shard (weak scaling) and the
    library's sharded join (ShardedJoin -> phj_dist_*) joins them; --workload scaled: BASELINE.json configs[4],
    160 M x 3.2 B in total (strong scaling). Before anything is timed a join with a non-trivial answer is checked
    against an independent count (`parity`)."""
    import json

    import torch

    import partitionedhashjoin_b200 as phj
    scaled = getattr(args, "workload", "default") == "scaled"
    mode = getattr(args, "shuffle", "library")
    dev = f"cuda:{local}"

    def make_job(kind=None):
        kind = kind or mode
        if kind in ("library", "library_no_hot"):
            from ._lib import FLAG_NO_HOT_DIGITS
            return ShardedJoin(dist if world > 1 else None, rank, world, local, partitions=args.dist_partitions,
                               chunks=args.chunks, hash=args.hash,
                               flags=FLAG_NO_HOT_DIGITS if kind == "library_no_hot" else 0)
        if kind in ("pass1", "fused"):
            return FusedShardedRadixJoin(dist, rank, world, FusedGpuBackend(
                world, local, partitions_local=args.partitions, hash=args.hash, pass1_in_shuffle=(kind == "pass1")))
        if kind == "npj":
            return ReplicatedNoPartitioningJoin(dist, rank, world, NpjGpuBackend(world, local, hash=args.hash))
        return ShardedRadixJoin(dist, rank, world, GpuBackend(world, local, partitions_local=args.partitions, hash=args.hash))

    if world == 1:
        dist = SingleRank
    # ---- parity first: a non-trivial count on uniform and on heavy-hitter keys; the legacy host-driven path with
    # ---- heavy-hitter digits kept local (layout_hot) on the same Zipf 1.25 data
    parity = parity_check(phj, dist, torch, rank, world, local, make_job, 0.01)
    parity_skew = parity_check(phj, dist, torch, rank, world, local, make_job, 1.25)
    legacy = world > 1 and not world & (world - 1)  # the round-1 host-driven paths want a power of two
    parity_hot = parity_check(phj, dist, torch, rank, world, local, lambda: make_job("pass1"), 1.25) if legacy else None
    # ... and the library path with PHJ_FLAG_NO_HOT_DIGITS (by default it keeps heavy-hitter digits local and replicates
    # their build side: `parity_zipf1.25` went that way, see its `hot_digits`)
    parity_lib_hot = parity_check(phj, dist, torch, rank, world, local, lambda: make_job("library_no_hot"), 1.25) \
        if world > 1 else None
    assert parity["ok"] and parity_skew["ok"] and (parity_hot is None or parity_hot["ok"]) and \
        (parity_lib_hot is None or parity_lib_hot["ok"]), (parity, parity_skew, parity_hot, parity_lib_hot)

    if scaled:
        dR, dS, n_build, n_probe, want = scaled_inputs(phj, dist, torch, rank, world, local, args.skew)
    else:
        n_build, n_probe = 10_000_000, 200_000_000
        Rp, Sp = shard_inputs(phj, rank, world, n_build, n_probe, args.skew, 12345, 64)
        want = world * n_probe  # generator data over R's own key range: every probe key has a build match
    job = make_job()
    if scaled:
        job.bind_device(dR.ptr, n_build, dS.ptr, n_probe, keepalive=(dR, dS))
    else:
        job.upload(Rp.array, Sp.array)
    # warm-up: CUDA events around EVERY kernel (the per-kernel table of the line); the timed steps keep them only
    # around radix_scatter, the roofline kernel -- as on one GPU (a pair of events costs its stream about a microsecond)
    warm_times = {}
    if hasattr(job, "kernel_timing"):
        job.kernel_timing("")
    for _ in range(args.warmup):
        res = job.join()
        if hasattr(job, "kernel_times"):
            for name, ns in job.kernel_times():
                warm_times.setdefault(name, []).append(ns)
    assert res["matches"] == want, (res["matches"], want)
    if hasattr(job, "kernel_timing"):
        job.kernel_timing("radix_scatter")

    def sync():
        torch.cuda.synchronize()
        dist.barrier()
        torch.cuda.synchronize()

    # Timed on the device: a CUDA event pair brackets the K joins (every join returns with its streams
    # synchronised, so the closing event's timestamp lies after the last kernel), barrier + synchronize on both
    # sides, max over ranks. The wall clock is kept beside it.
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches0 = job.launches if hasattr(job, "launches") else job.backend.launches
    keys = ("count_ns", "shuffle_ns", "build_ns", "probe_ns", "total_ns")
    sums, ktimes = {k: 0 for k in keys}, {}
    with ClockSampler(local) as clocks:
        sync()
        ev0.record()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            res = job.join()
            for k in keys:
                sums[k] += res.get(k, 0)
            if hasattr(job, "kernel_times"):
                for name, ns in job.kernel_times():
                    ktimes.setdefault(name, []).append(ns)
        ev1.record()
        sync()
        wall = time.perf_counter() - t0
    assert res["matches"] == want
    elapsed, wall = _allreduce(dist, torch, local, [ev0.elapsed_time(ev1) / 1e3, wall], dist.ReduceOp.MAX)
    launches1 = job.launches if hasattr(job, "launches") else job.backend.launches
    launches = int(_allreduce(dist, torch, local, [float(launches1 - launches0)])[0])

    # NVLink reference, measured in this run: a plain peer copy, every rank into its neighbour's window at once
    peer_gbps = job.measure_peer_copy() if hasattr(job, "measure_peer_copy") else 0.0
    peer_gbps = _allreduce(dist, torch, local, [peer_gbps], dist.ReduceOp.MIN)[0]
    sent = _allreduce(dist, torch, local, [float(res.get("shuffle_bytes", 0))], dist.ReduceOp.MAX)[0]

    # e2e: the host shards are uploaded inside the timed region (the scaled workload is generated on the device
    # and has no host copy, hence no end-to-end number)
    e2e_steps, e2e_s = max(3, min(args.steps, 5)), 0.0
    if not scaled:
        sync()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            job.upload(Rp.array, Sp.array)
            r2 = job.join()
        sync()
        e2e_s = _allreduce(dist, torch, local, [time.perf_counter() - t0], dist.ReduceOp.MAX)[0]
        assert r2["matches"] == want
    lay = job.info() if hasattr(job, "info") else {}
    job.close()

    # ---- the other multi-GPU configurations, a few joins each ----
    others = {}
    if not scaled and not args.quick:
        def few(j, n=3, expect=want):
            j.upload(Rp.array, Sp.array)
            runs = [j.join() for _ in range(n + 1)][1:]
            for r in runs:
                assert r["matches"] == expect, (r["matches"], expect)
            sync()
            t0 = time.perf_counter()
            for _ in range(n):
                j.join()
            sync()
            dt = _allreduce(dist, torch, local, [time.perf_counter() - t0], dist.ReduceOp.MAX)[0] / n
            j.close()
            return round(world * (n_build + n_probe) / dt / 1e9, 2)
        if world > 1:
            others["no_partitioning_replicated_build_Gtuples_s"] = few(make_job("npj"))
        if legacy:
            others["legacy_host_driven_pass1_Gtuples_s"] = few(make_job("pass1"))
        others["library_1_chunk_no_overlap_Gtuples_s"] = few(ShardedJoin(dist, rank, world, local, chunks=1, hash=args.hash))
        for skew in (1.05, 1.25):
            phj.fill_zipf(Sp.array, skew, 1, world * n_build, 12345 + 7919 * rank, 64)
            others[f"library_zipf{skew}_Gtuples_s"] = few(make_job("library"))
            if skew == 1.25 and world > 1:
                others["library_no_hot_digits_zipf1.25_Gtuples_s"] = few(make_job("library_no_hot"))
                if legacy:
                    others["legacy_pass1_hot_digits_local_zipf1.25_Gtuples_s"] = few(make_job("pass1"))
        Rp.close()
        Sp.close()
        others["scaled_160Mx3200M"] = scaled_other_config(phj, dist, torch, rank, world, local, args)

    if rank == 0:
        n_tuples = world * (n_build + n_probe)
        ms = elapsed / args.steps * 1e3
        peak, peak_src = measured_hbm_peak()
        cfg = workload_config(args)
        cfg["primary"], cfg["secondary"] = world * n_build, world * n_probe
        cfg["partitions"] = lay.get("digits", args.partitions)
        cfg["l2_flush"] = f"inputs ({16 * (n_build + n_probe) / 1e9:.2f} GB per GPU) are larger than L2"
        if mode == "library":
            cfg["workload"] = (
                f"radix join sharded over {world} B200: {world} x ({n_build // 10**6}M x {n_probe // 10**6}M) row shards = "
                f"{world * n_build // 10**6}M x {world * n_probe // 10**6}M{' (device-generated, keys over 2x the build range)' if scaled else ''}; "
                f"ONE partitioning pass = the shuffle ({lay.get('digits')} digits = {world} GPUs x {lay.get('local_partitions')} local "
                f"partitions, NVLink peer stores from the scatter kernel, {lay.get('chunks')} probe chunks overlapping the local "
                f"L2-table probes), NCCL for sizes / barriers / count, {args.hash}, Zipf skew {args.skew}")
        else:
            cfg["workload"] = f"radix join sharded over {world} B200, shuffle mode {mode}, {args.hash}, Zipf skew {args.skew}"
        cfg["parallelism"] = f"partition-sharded x{world}"
        scat = [ns for name, v in ktimes.items() if name.startswith("radix_scatter") for ns in v]
        scat_ms = sum(scat) / max(args.steps, 1) / 1e6  # all scatter launches of one join, this rank
        alg_bytes = 32.0 * (n_build + n_probe)
        shuffle_ms = sums["shuffle_ns"] / args.steps / 1e6
        line = {
            "metric": metric, "value": n_tuples / (elapsed / args.steps), "unit": unit, "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
            "timing": "CUDA events around the K joins, max over ranks", "wall_ms_per_step": wall / args.steps * 1e3,
            "scaling": "strong" if scaled else "weak", "vs_baseline": None, "dtype": "int64", "data": "synthetic",
            "config": cfg,
            "e2e": None if scaled else {
                "value": n_tuples / (e2e_s / e2e_steps), "unit": unit, "h2d_bytes_per_step": 16 * n_tuples,
                "d2h_bytes_per_step": world * (64 + 8 * world * 2 * lay.get("digits", 64) * (lay.get("chunks", 1) + 1)),
                "steps": e2e_steps, "host_memory": "pinned (phj_host_alloc)"},
            "gpu_launches": launches,
            "parity": parity, "parity_zipf1.25": parity_skew, "parity_zipf1.25_legacy_hot_digits": parity_hot,
            "parity_zipf1.25_library_no_hot_digits": parity_lib_hot,
            "roofline": {"bound": "hbm", "kernel": "radix_scatter[shuffle] (rank 0, all launches of a join; NVLink-bound "
                                                   "for N > 1: see `shuffle`)", "unit": "GB/s", "peak": peak,
                         "peak_source": peak_src, "achieved": alg_bytes / max(scat_ms, 1e-9) / 1e6,
                         "frac": alg_bytes / max(scat_ms, 1e-9) / 1e6 / peak,
                         "traffic": shuffle_scatter_dram_traffic(world)[0] if mode == "library" and not scaled else None,
                         "traffic_source": shuffle_scatter_dram_traffic(world)[1] if mode == "library" and not scaled else None,
                         "algorithmic_bytes_per_join_per_gpu": alg_bytes, "ms_per_join": scat_ms},
            "phases_ms_rank0": {k[:-3]: v / args.steps / 1e6 for k, v in sums.items()},
            # medians over the warm-up joins (which time every kernel) and, for the scatter, the timed steps
            "kernel_us_rank0": {name: round(sorted(v)[len(v) // 2] / 1e3, 1) for name, v in {**warm_times, **ktimes}.items()},
            "shuffle": {"bytes_sent_remote_per_gpu": sent, "ms": shuffle_ms,
                        "GBps_per_gpu": sent / max(shuffle_ms, 1e-9) / 1e6,
                        "nvlink_peak_GBps": peer_gbps,
                        "nvlink_peak_source": "plain peer cudaMemcpyAsync of 1 GiB into the next rank's window, all "
                                              "ranks at once, measured in this run",
                        "nvlink_frac": (sent / max(shuffle_ms, 1e-9) / 1e6) / peer_gbps if peer_gbps else None,
                        "nvlink_nominal_GBps": 900.0},
            "layout": lay, "clocks": clocks.summary(), "matches": res["matches"], "want": want,
            "cpu_baseline": None, "other_configs": others,
        }
        print(json.dumps(line), flush=True)
    dist.barrier()
    dist.destroy_process_group()
