"""Profiling driver for the NVLink scatter of the sharded join: ONE process, two GPUs, no NCCL and no cross-GPU
synchronisation, so it is safe under `ncu --set full` (kernel replay) with `--devices 0`:

    GPU 0 holds one rank's row shard (10 M x 200 M, device-generated) and plays rank 0 of `world` ranks;
    GPU 1 holds ONE window that stands in for the windows of all the other ranks.

It runs the host-driven half-steps of the multi-GPU path (phj_shard_count, phj_shard_scatter): the same
radix_scatter<6, ..., 1024, 8> launch as the library's sharded join, every digit run of a remote owner stored
over NVLink with TMA bulk stores, the own digits into a local window.

    python tools/ncu_split.py [world] [digits] [scatters]
"""
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("PHJ_KERNEL_TIMES", "1")
import partitionedhashjoin_b200 as phj  # noqa: E402
from partitionedhashjoin_b200 import _lib  # noqa: E402


def main():
    world = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    ndig = int(sys.argv[2]) if len(sys.argv) > 2 else 64
    reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
    assert phj.device_count() >= 2, "needs two GPUs"
    lib, check = _lib.lib, _lib.check
    check(lib.phj_enable_peer_access(0, 1))
    nr, ns = 10_000_000, 200_000_000
    dR = phj.DeviceTuples(nr, 0).fill_sequential(1)
    dS = phj.DeviceTuples(ns, 0).fill_zipf(0.01, 1, world * nr, 12345, 1 << 14)
    with phj.Engine("shard-split", partitions=ndig, device=0, shard_shift=0) as e:
        e.bind_device(dR.ptr, nr, dS.ptr, ns, keepalive=(dR, dS))
        counts = np.zeros((2, ndig), dtype=np.uint64)
        check(lib.phj_shard_count(e._h, counts.ctypes.data))
        counts = counts.astype(np.int64)
        dl = ndig // world
        own = np.arange(ndig) // dl == 0          # rank 0 owns the first ndig / world digits
        excl = np.cumsum(counts, axis=1) - counts
        # two windows per relation: the local one (own digits) and the remote one (everything else), digit-major
        wins = {}
        for rel in (0, 1):
            for dev, mask in ((0, own), (1, ~own)):
                rows = int(counts[rel][mask].sum()) + 4096
                ptr, handle = C.c_void_p(), (C.c_ubyte * 64)()
                check(lib.phj_shared_alloc(dev, rows * 16, C.byref(ptr), handle))
                wins[rel, dev] = ptr.value
        args = []
        for rel in (0, 1):
            off = np.zeros(ndig, dtype=np.uint64)
            for mask in (own, ~own):
                c = counts[rel] * mask
                off[mask] = (np.cumsum(c) - c)[mask]
            ptrs = (C.c_void_p * ndig)(*[C.c_void_p(wins[rel, 0 if own[d] else 1]) for d in range(ndig)])
            args += [ptrs, np.ascontiguousarray(off)]
        remote = int(counts[:, ~own].sum()) * 16
        res = _lib.PhjResult()
        for it in range(reps):
            if it:  # a scatter consumes its count
                again = np.zeros((2, ndig), dtype=np.uint64)
                check(lib.phj_shard_count(e._h, again.ctypes.data))
            check(lib.phj_shard_scatter(e._h, 0, args[0], args[1].ctypes.data, args[2], args[3].ctypes.data, C.byref(res)))
            kt = dict(e.kernel_times())
            ns_scatter = sum(t for n, t in e.kernel_times() if "scatter" in n)
            print(f"world {world} digits {ndig}: scatter {ns_scatter / 1e6:.3f} ms, {remote / 1e9:.2f} GB over NVLink = "
                  f"{remote / max(ns_scatter, 1):.0f} GB/s; kernels {[(n, round(t / 1e3, 1)) for n, t in e.kernel_times()]}",
                  flush=True)
        _ = excl, kt
    for (rel, dev), p in wins.items():
        check(lib.phj_shared_free(dev, C.c_void_p(p)))


if __name__ == "__main__":
    main()
