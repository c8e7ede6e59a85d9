"""Python host mirror of the reference's joiner interface, on top of the C ABI (include/phj.h).

The C++ mirror (what a reference maintainer would use) lives in ``host/``; this module gives the
tests and ``bench.py`` the same surface with the reference's names:

* ``NoPartitioningHashJoiner(configuration).Run(tableA, tableB, timer)``
  -- ``NoPartitioning::HashJoiner::Run`` (reference src/NoPartitioning/HashJoin.hpp:54-74)
* ``RadixClusteringHashJoiner(configuration, hasher).Run(tableA, tableB, timer)``
  -- ``RadixClustering::HashJoiner::Run`` (reference src/RadixCluster/HashJoin.hpp:190-241)
* ``HashJoinTimer`` / ``NoOpHashJoinTimer`` -- ``Common::IHashJoinTimer`` (src/Common/Results.hpp:131-247)

Everything that computes goes through ``libphj_b200.so``; nothing here falls back to numpy.
"""
from __future__ import annotations

import ctypes as C
import dataclasses
from typing import Optional

import numpy as np

from . import _lib
from ._lib import (ALGO_NO_PARTITIONING, ALGO_RADIX_PARTITIONING, HASH_NAMES, TUPLE_DTYPE, PhjConfig,
                   PhjDeviceInfo, PhjError, PhjResult, check, lib)


# ---------------------------------------------------------------------------------------------
# host data helpers
# ---------------------------------------------------------------------------------------------
def as_tuples(a) -> np.ndarray:
    """View/convert to a C-contiguous array of TUPLE_DTYPE (16-byte {id, payload} records)."""
    a = np.asarray(a)
    if a.dtype != TUPLE_DTYPE:
        if a.dtype.fields is None and a.ndim == 2 and a.shape[1] == 2:
            a = np.ascontiguousarray(a, dtype="<i8").view(TUPLE_DTYPE).reshape(-1)
        else:
            raise TypeError(f"expected TUPLE_DTYPE or an (n, 2) int64 array, got {a.dtype} {a.shape}")
    return np.ascontiguousarray(a)


def make_tuples(ids, payloads=None) -> np.ndarray:
    ids = np.asarray(ids, dtype=np.int64)
    out = np.empty(ids.shape[0], dtype=TUPLE_DTYPE)
    out["id"] = ids
    out["payload"] = np.arange(ids.shape[0], dtype=np.int64) if payloads is None else payloads
    return out


class PinnedTuples:
    """A relation in page-locked host memory (phj_host_alloc): what phj_join_host uploads from."""

    def __init__(self, n: int):
        self.n = int(n)
        self._ptr = C.c_void_p()
        check(lib.phj_host_alloc(C.byref(self._ptr), max(self.n, 1) * 16))
        buf = (C.c_char * (max(self.n, 1) * 16)).from_address(self._ptr.value)
        self.array = np.frombuffer(buf, dtype=TUPLE_DTYPE, count=self.n)

    def close(self):
        if self._ptr:
            self.array = None
            lib.phj_host_free(self._ptr)
            self._ptr = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def fill_sequential(out: np.ndarray, start: int = 1, threads: int = 0) -> np.ndarray:
    """DataGenerator::Sequential::FillTable (reference src/DataGenerator/Sequential.cpp:6-40)."""
    assert out.dtype == TUPLE_DTYPE and out.flags.c_contiguous
    check(lib.phj_fill_sequential(out.ctypes.data, out.shape[0], start, threads))
    return out


def fill_zipf(out: np.ndarray, alpha: float, range_first: int, range_second: int, base_seed: int,
              batches: int = 64, threads: int = 0) -> np.ndarray:
    """DataGenerator::Zipf::FillTable with explicit seeding (reference src/DataGenerator/Zipf.cpp:58-108)."""
    assert out.dtype == TUPLE_DTYPE and out.flags.c_contiguous
    rc = lib.phj_fill_zipf(out.ctypes.data, out.shape[0], alpha, range_first, range_second, base_seed,
                           batches, threads)
    if rc != 0:
        raise ValueError("phj_fill_zipf rejected its arguments (alpha < 0.01, empty range, 0 batches "
                         "or a seed that is a multiple of 2^31-1)")
    return out


class DeviceTuples:
    """A relation in device memory (phj_shared_alloc): target of the device-side generators, source
    of Engine.bind_device, mappable by peer processes."""

    def __init__(self, n: int, device: int = 0):
        self.n, self.device = int(n), device
        self._ptr, self.ipc_handle = C.c_void_p(), (C.c_ubyte * 64)()
        check(lib.phj_shared_alloc(device, max(self.n, 1) * 16, C.byref(self._ptr), self.ipc_handle))

    @property
    def ptr(self) -> int:
        return self._ptr.value or 0

    def fill_sequential(self, start: int = 1):
        check(lib.phj_device_fill_sequential(self.device, self._ptr, self.n, start))
        return self

    def fill_zipf(self, alpha: float, range_first: int, range_second: int, base_seed: int, batches: int):
        check(lib.phj_device_fill_zipf(self.device, self._ptr, self.n, alpha, range_first, range_second, base_seed,
                                       batches))
        return self

    def upload(self, rel: np.ndarray):
        rel = as_tuples(rel)
        assert rel.shape[0] == self.n
        check(lib.phj_memcpy_h2d(self.device, self._ptr, rel.ctypes.data, self.n * 16))
        return self

    def download(self, first: int = 0, count: Optional[int] = None) -> np.ndarray:
        count = self.n - first if count is None else count
        out = np.empty(count, dtype=TUPLE_DTYPE)
        check(lib.phj_memcpy_d2h(self.device, out.ctypes.data, C.c_void_p(self.ptr + first * 16), count * 16))
        return out

    def close(self):
        if self._ptr:
            lib.phj_shared_free(self.device, self._ptr)
            self._ptr = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def hash_host(hash_id: int, seed: int, key: int) -> int:
    return int(lib.phj_hash_host(hash_id, C.c_uint64(seed & (2**64 - 1)), key))


def hash_batch(hash_id: int, seed: int, keys: np.ndarray, device: int = 0) -> np.ndarray:
    """Raw 64-bit hashes computed ON THE DEVICE (test hook for the __device__ hashers)."""
    keys = np.ascontiguousarray(keys, dtype=np.int64)
    out = np.empty(keys.shape[0], dtype=np.uint64)
    check(lib.phj_hash_batch(hash_id, C.c_uint64(seed & (2**64 - 1)), keys.ctypes.data, keys.shape[0],
                             out.ctypes.data, device))
    return out


def device_count() -> int:
    return int(lib.phj_device_count())


def device_info(device: int = 0) -> dict:
    info = PhjDeviceInfo()
    check(lib.phj_get_device_info(device, C.byref(info)))
    d = {name: getattr(info, name) for name, _ in info._fields_}
    d["name"] = info.name.decode()
    return d


# ---------------------------------------------------------------------------------------------
# thin handle wrapper
# ---------------------------------------------------------------------------------------------
class Engine:
    """One phj_handle. ``algo``: 'no-partitioning' | 'radix-partitioning' (the CLI spellings,
    reference src/Common/Configuration.cpp:4-12)."""

    ALGOS = {"no-partitioning": ALGO_NO_PARTITIONING, "radix-partitioning": ALGO_RADIX_PARTITIONING,
             "shard-split": _lib.ALGO_SHARD_SPLIT}

    def __init__(self, algo="radix-partitioning", partitions: int = 0, radix_bits=(0, 0), hash="xxh3",
                 hash_seed: int = 0x9E3779B97F4A7C15, table_seed: int = 1, device: int = 0,
                 flags: int = 0, shard_shift: int = 0, split_ctas: int = 0, reserve=(0, 0),
                 split_chunks: int = 0, upload_chunks: int = 0, num_gpus: int = 0):
        cfg = PhjConfig()
        if isinstance(algo, str):
            if algo not in self.ALGOS:
                raise ValueError(f"Unrecognized join algorithm type: {algo}.")
            algo = self.ALGOS[algo]
        cfg.algo = algo
        cfg.hash = HASH_NAMES[hash] if isinstance(hash, str) else hash
        cfg.partitions = partitions
        cfg.radix_bits[0], cfg.radix_bits[1] = radix_bits
        cfg.hash_seed = hash_seed & (2**64 - 1)
        cfg.table_seed = table_seed & (2**64 - 1)
        cfg.device = device
        cfg.flags = flags
        cfg.shard_shift = shard_shift
        cfg.split_ctas = split_ctas
        cfg.split_chunks = split_chunks
        cfg.upload_chunks = upload_chunks  # join_host: 0 = automatic, 1 = upload then join, k = k probe chunks
        cfg.reserve_build, cfg.reserve_probe = reserve
        # num_gpus > 1: this process drives GPUs device .. device + num_gpus - 1 (the sharded join of
        # csrc/phj_dist.inl); partitions then = GPUs x local partitions (<= 256), split_chunks = probe chunks
        cfg.num_gpus = num_gpus
        self._h = C.c_void_p()
        check(lib.phj_create(C.byref(cfg), C.byref(self._h)))
        self._keep = []

    def close(self):
        if getattr(self, "_h", None):
            lib.phj_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def upload(self, build: np.ndarray, probe: np.ndarray):
        build, probe = as_tuples(build), as_tuples(probe)
        check(lib.phj_upload(self._h, build.ctypes.data, build.shape[0], probe.ctypes.data, probe.shape[0]))
        self._n = (build.shape[0], probe.shape[0])

    def bind_device(self, d_build: int, n_build: int, d_probe: int, n_probe: int, keepalive=None):
        """Join relations already resident on the device (raw device pointers, e.g. tensor.data_ptr())."""
        check(lib.phj_bind_device(self._h, C.c_void_p(d_build), n_build, C.c_void_p(d_probe), n_probe))
        self._keep = [keepalive]
        self._n = (n_build, n_probe)

    def bind_device_partitioned(self, d_build: int, n_build: int, d_probe: int, n_probe: int, bounds_build,
                                bounds_probe, keepalive=None, first_parent: int = 0, parent_space: int = 0,
                                parent_digits=None):
        """Device-resident relations already partitioned by pass-1 digit; bounds_* are the parents + 1
        host boundaries of each relation. The parents are digits [first_parent, first_parent + parents)
        of a digit space of `parent_space` values (0 = this plan's own 2^b1)."""
        bb = np.ascontiguousarray(bounds_build, dtype=np.uint64)
        bp = np.ascontiguousarray(bounds_probe, dtype=np.uint64)
        if bb.shape != bp.shape or bb.ndim != 1 or bb.shape[0] < 2:
            raise ValueError("bounds_build and bounds_probe must both hold parents + 1 boundaries")
        check(lib.phj_bind_device_partitioned(self._h, C.c_void_p(d_build), n_build, C.c_void_p(d_probe), n_probe,
                                              bb.ctypes.data, bp.ctypes.data, bb.shape[0] - 1, first_parent,
                                              parent_space))
        if parent_digits is not None:  # parents that are not the contiguous range first_parent ..
            pd = np.ascontiguousarray(parent_digits, dtype=np.uint32)
            check(lib.phj_set_parent_digits(self._h, pd.ctypes.data, pd.shape[0]))
        self._keep = [keepalive]
        self._n = (n_build, n_probe)

    def join(self) -> dict:
        res = PhjResult()
        check(lib.phj_join(self._h, C.byref(res)))
        return res.as_dict()

    def join_materialize(self) -> dict:
        """The join with its result: res['joined_tuples'] rows wait on the device (read_joined)."""
        res = PhjResult()
        check(lib.phj_join_materialize(self._h, C.byref(res)))
        self._joined = int(res.joined_tuples)
        return res.as_dict()

    def read_joined(self, first: int = 0, count: Optional[int] = None) -> np.ndarray:
        """Rows [first, first + count) of the last join_materialize as a JOINED_DTYPE array."""
        count = self._joined - first if count is None else count
        out = np.empty(count, dtype=_lib.JOINED_DTYPE)
        check(lib.phj_read_joined(self._h, out.ctypes.data, first, count))
        return out

    def join_host(self, build: np.ndarray, probe: np.ndarray) -> dict:
        """Host relations in, count out (phj_join_host): the uploads are part of the call, and for a
        large probe relation they overlap the joins of the chunks that have already landed."""
        build, probe = as_tuples(build), as_tuples(probe)
        res = PhjResult()
        check(lib.phj_join_host(self._h, build.ctypes.data, build.shape[0], probe.ctypes.data,
                                probe.shape[0], C.byref(res)))
        self._n = (build.shape[0], probe.shape[0])
        return res.as_dict()

    def read_partitions(self, which: int, partitions: int):
        """(partitioned relation, partitions+1 boundaries) of the last radix join."""
        n = self._n[which]
        out = np.empty(n, dtype=TUPLE_DTYPE)
        bounds = np.empty(partitions + 1, dtype=np.uint64)
        check(lib.phj_read_partitions(self._h, which, out.ctypes.data, bounds.ctypes.data))
        return out, bounds

    def read_bounds(self, which: int, partitions: int) -> np.ndarray:
        """Only the partitions+1 boundaries of the last radix join / shard split."""
        bounds = np.empty(partitions + 1, dtype=np.uint64)
        check(lib.phj_read_partitions(self._h, which, None, bounds.ctypes.data))
        return bounds

    def device_partitions(self, which: int):
        """(device pointer of the partitioned relation, device pointer of its boundaries, n)."""
        d_data, d_bounds, n = C.c_void_p(), C.c_void_p(), C.c_size_t()
        check(lib.phj_device_partitions(self._h, which, C.byref(d_data), C.byref(d_bounds), C.byref(n)))
        return d_data.value or 0, d_bounds.value or 0, int(n.value)

    def kernel_timing(self, filter) -> None:
        """CUDA events around the kernels of the following joins: None = none, "" = all, else names containing it."""
        check(lib.phj_kernel_timing(self._h, None if filter is None else filter.encode()))

    def kernel_times(self) -> list:
        names = (C.c_char_p * 32)()
        ns = (C.c_uint64 * 32)()
        k = lib.phj_kernel_times(self._h, names, ns, 32)
        return [(names[i].decode(), int(ns[i])) for i in range(k)]


# ---------------------------------------------------------------------------------------------
# the reference's interface, by its own names
# ---------------------------------------------------------------------------------------------
@dataclasses.dataclass
class NoPartitioningConfiguration:
    """NoPartitioning::Configuration (reference src/NoPartitioning/Configuration.hpp:6-8)."""
    MinBatchSize: int = 10000  # CPU batching knob; accepted for source compatibility, unused


@dataclasses.dataclass
class RadixClusteringConfiguration:
    """RadixClustering::Configuration (reference src/RadixCluster/Configuration.hpp:6-9)."""
    MinBatchSize: int = 10000
    NumberOfPartitions: int = 32


@dataclasses.dataclass
class Hasher:
    """Stands in for a Common::IHasher instance: which function and which seed."""
    name: str = "xxh3"
    seed: int = 0x9E3779B97F4A7C15


class HashJoinTimingResult:
    """Common::HashJoinTimingResult (reference src/Common/Results.hpp:60-88); nanoseconds."""

    def __init__(self, build=0, probe=0, partitioning=0, parameters=None):
        self.build_ns, self.probe_ns, self.partitioning_ns = build, probe, partitioning
        self.parameters = dict(parameters or {})


class NoOpHashJoinTimer:
    """Common::NoOpHashJoinTimer (reference src/Common/Results.hpp:151-165)."""

    def SetBuildPhaseDuration(self, ns): pass
    def SetProbePhaseDuration(self, ns): pass
    def SetPartitionPhaseDuration(self, ns): pass
    def GetResult(self): return HashJoinTimingResult()


class HashJoinTimer(NoOpHashJoinTimer):
    """Common::HashJoinTimer's set-duration interface (reference src/Common/Results.hpp:213-240).
    The device path reports measured durations, so only the thread-safe Set*Duration half of
    IHashJoinTimer is driven."""

    def __init__(self, parameters=None):
        self._r = HashJoinTimingResult(parameters=parameters)

    def SetBuildPhaseDuration(self, ns): self._r.build_ns = int(ns)
    def SetProbePhaseDuration(self, ns): self._r.probe_ns = int(ns)
    def SetPartitionPhaseDuration(self, ns): self._r.partitioning_ns = int(ns)
    def GetResult(self): return self._r


class _JoinerBase:
    def __init__(self, engine: Engine, materialize: bool = False, stream_upload: bool = False):
        self._engine = engine
        self._materialize = materialize
        self._stream_upload = stream_upload and not materialize
        self.last_result: Optional[dict] = None

    def Run(self, tableA, tableB, timer=None):
        """tableA is the build relation, tableB the probe relation. Like the reference the returned
        joined table is empty (count-only join, reference Readme.md:10); the count is in
        ``self.last_result['matches']``."""
        timer = timer or NoOpHashJoinTimer()
        if self._stream_upload:  # one call; the upload of tableB overlaps the joins of its landed chunks
            res = self._engine.join_host(tableA, tableB)
        else:
            self._engine.upload(tableA, tableB)
            res = self._engine.join_materialize() if self._materialize else self._engine.join()
        self.last_result = res
        timer.SetPartitionPhaseDuration(res["partition_ns"])
        timer.SetBuildPhaseDuration(res["build_ns"])
        timer.SetProbePhaseDuration(res["probe_ns"])
        if self._materialize:  # the Table<JoinedTuple> the reference declares but leaves empty
            return self._engine.read_joined()
        return np.empty(0, dtype=_lib.JOINED_DTYPE)

    def close(self):
        self._engine.close()


class NoPartitioningHashJoiner(_JoinerBase):
    """``table``: 'linear-probing' (HashTables::LinearProbingHashTable, what the reference's main
    instantiates, src/main.cpp:216-217) or 'separate-chaining' (HashTables::SeparateChainingHashTable,
    src/HashTables/SeparateChaining.hpp) -- the HashTableFactory template argument of the reference."""

    TABLES = {"linear-probing": 0, "separate-chaining": _lib.FLAG_CHAINED_TABLE}

    def __init__(self, configuration: Optional[NoPartitioningConfiguration] = None,
                 hasher: Optional[Hasher] = None, device: int = 0, table: str = "linear-probing",
                 stream_upload: bool = False):
        hasher = hasher or Hasher()
        if table not in self.TABLES:
            raise ValueError(f"Unrecognized hash table type: {table}.")
        super().__init__(Engine("no-partitioning", hash=hasher.name, hash_seed=hasher.seed, device=device,
                                flags=self.TABLES[table]), stream_upload=stream_upload)
        self.configuration = configuration or NoPartitioningConfiguration()


class RadixClusteringHashJoiner(_JoinerBase):
    def __init__(self, configuration: Optional[RadixClusteringConfiguration] = None,
                 hasher: Optional[Hasher] = None, device: int = 0, radix_bits=(0, 0), flags: int = 0,
                 materialize: bool = False, stream_upload: bool = False):
        configuration = configuration or RadixClusteringConfiguration()
        hasher = hasher or Hasher()
        if configuration.NumberOfPartitions < 0:
            raise ValueError("NumberOfPartitions must be >= 0")
        super().__init__(Engine("radix-partitioning", partitions=configuration.NumberOfPartitions,
                                radix_bits=radix_bits, hash=hasher.name, hash_seed=hasher.seed,
                                device=device, flags=flags), materialize=materialize, stream_upload=stream_upload)
        self.configuration = configuration
