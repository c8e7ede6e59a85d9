"""ctypes access to the two CPU checkers (TEST INFRASTRUCTURE):

* ``oracle/libphj_oracle.so``  -- the plain-C restatement (oracle/phj_oracle.c), built on demand;
* ``oracle/_ref/libphj_ref.so`` -- the UNMODIFIED reference compiled with shims (oracle/ref_harness.cpp);
  built in the authoring container (needs /root/reference), prebuilt on the GPU box.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
ORACLE_SO = os.path.join(ORACLE_DIR, "libphj_oracle.so")
REF_SO = os.path.join(ORACLE_DIR, "_ref", "libphj_ref.so")
REFERENCE_SRC = "/root/reference/src"

TUPLE = np.dtype([("id", "<i8"), ("payload", "<i8")], align=True)
u64, i64, sz, dbl, vp = C.c_uint64, C.c_int64, C.c_size_t, C.c_double, C.c_void_p


JOINED = np.dtype([("id", "<i8"), ("payloadA", "<i8"), ("payloadB", "<i8")])


class OResult(C.Structure):
    _fields_ = [("matches", u64), ("partition_ns", u64), ("build_ns", u64), ("probe_ns", u64)]


class RefResult(C.Structure):
    _fields_ = [("matches", u64), ("partition_ns", u64), ("build_ns", u64), ("probe_ns", u64),
                ("wall_ns", u64), ("workers", C.c_uint32), ("joined_records", C.c_uint32)]


def build_oracle(force=False):
    src = [os.path.join(ORACLE_DIR, f) for f in ("phj_oracle.c", "phj_oracle.h")]
    stale = (not os.path.exists(ORACLE_SO)) or any(os.path.getmtime(s) > os.path.getmtime(ORACLE_SO) for s in src)
    if force or stale:
        subprocess.run(["make", "-C", ORACLE_DIR, "oracle"], check=True, capture_output=True)
    return ORACLE_SO


def build_ref(force=False):
    """Compile the reference where it lies; a no-op (keeps the prebuilt .so) when it is absent."""
    if not os.path.isdir(REFERENCE_SRC):
        return REF_SO if os.path.exists(REF_SO) else None
    srcs = [os.path.join(ORACLE_DIR, "ref_harness.cpp")]
    stale = (not os.path.exists(REF_SO)) or any(os.path.getmtime(s) > os.path.getmtime(REF_SO) for s in srcs)
    if force or stale:
        subprocess.run(["make", "-C", ORACLE_DIR, "ref"], check=True, capture_output=True)
    return REF_SO


def _sig(lib, name, restype, argtypes):
    fn = getattr(lib, name)
    fn.restype, fn.argtypes = restype, argtypes
    return fn


class Oracle:
    """The C restatement."""

    def __init__(self):
        self.lib = lib = C.CDLL(build_oracle())
        _sig(lib, "phjo_xxh3_64", u64, [i64, u64])
        _sig(lib, "phjo_murmur3_64", u64, [i64, u64])
        _sig(lib, "phjo_city_64", u64, [i64, u64])
        _sig(lib, "phjo_hash", u64, [C.c_int, i64, u64, u64])
        _sig(lib, "phjo_hash_raw", u64, [C.c_int, i64, u64])
        _sig(lib, "phjo_hash_batch", None, [C.c_int, u64, vp, sz, vp])
        _sig(lib, "phjo_lcg_next", dbl, [C.POINTER(C.c_long)])
        _sig(lib, "phjo_zipf_generate", C.c_int, [dbl, u64, C.POINTER(C.c_long), C.POINTER(u64)])
        _sig(lib, "phjo_fill_sequential", None, [vp, sz, i64])
        _sig(lib, "phjo_fill_zipf_seeded", C.c_int, [vp, sz, dbl, i64, i64, C.c_long, sz])
        _sig(lib, "phjo_table_new", vp, [C.c_int, dbl, C.c_int, u64, sz])
        _sig(lib, "phjo_table_free", None, [vp])
        _sig(lib, "phjo_table_buckets", u64, [vp])
        _sig(lib, "phjo_table_insert", C.c_int, [vp, i64, vp])
        _sig(lib, "phjo_table_exists", C.c_int, [vp, i64])
        _sig(lib, "phjo_table_get", vp, [vp, i64])
        _sig(lib, "phjo_table_get_all", sz, [vp, i64, vp, sz])
        _sig(lib, "phjo_join_no_partitioning", C.c_int, [vp, sz, vp, sz, C.c_int, C.c_int, u64, C.POINTER(OResult)])
        _sig(lib, "phjo_partitioning_configuration", None,
             [sz, sz, sz, sz, C.POINTER(sz), C.POINTER(sz), C.POINTER(sz)])
        _sig(lib, "phjo_radix_partition", None, [vp, sz, sz, C.c_int, u64, sz, sz, vp, vp, vp])
        _sig(lib, "phjo_join_radix", C.c_int, [vp, sz, vp, sz, sz, sz, C.c_int, C.c_int, u64, u64, C.POINTER(OResult)])
        _sig(lib, "phjo_count_by_sort", u64, [vp, sz, vp, sz])
        _sig(lib, "phjo_join_materialize", u64, [vp, sz, vp, sz, C.c_int, C.c_int, u64, vp, u64])

    # hashing
    def hash_raw(self, hash_id, key, seed):
        return int(self.lib.phjo_hash_raw(hash_id, int(key), u64(int(seed) & (2**64 - 1))))

    def hash_batch(self, hash_id, seed, keys):
        keys = np.ascontiguousarray(keys, dtype=np.int64)
        out = np.empty(keys.shape[0], dtype=np.uint64)
        self.lib.phjo_hash_batch(hash_id, u64(int(seed) & (2**64 - 1)), keys.ctypes.data, keys.shape[0], out.ctypes.data)
        return out

    # generators
    def lcg_doubles(self, seed, n):
        st = C.c_long(seed)
        return [self.lib.phjo_lcg_next(C.byref(st)) for _ in range(n)]

    def zipf_samples(self, alpha, cardinality, seed, n):
        st, s = C.c_long(seed), u64()
        out = []
        for _ in range(n):
            if self.lib.phjo_zipf_generate(alpha, cardinality, C.byref(st), C.byref(s)):
                raise ValueError("alpha < 0.01")
            out.append(int(s.value))
        return out

    def fill_sequential(self, n, start=1):
        a = np.empty(n, dtype=TUPLE)
        self.lib.phjo_fill_sequential(a.ctypes.data, n, start)
        return a

    def fill_zipf(self, n, alpha, lo, hi, base_seed, batches):
        a = np.empty(n, dtype=TUPLE)
        if self.lib.phjo_fill_zipf_seeded(a.ctypes.data, n, alpha, lo, hi, base_seed, batches):
            raise ValueError("bad Zipf arguments")
        return a

    # joins
    def join_npj(self, R, S, table_kind=0, hash_id=0, seed_table=1):
        R, S = np.ascontiguousarray(R), np.ascontiguousarray(S)
        res = OResult()
        rc = self.lib.phjo_join_no_partitioning(R.ctypes.data, R.shape[0], S.ctypes.data, S.shape[0],
                                                table_kind, hash_id, u64(seed_table), C.byref(res))
        if rc == 1:
            raise ValueError("numberOfObjects must be greater than zero")
        assert rc == 0
        return int(res.matches)

    def join_radix(self, R, S, P, workers=7, table_kind=0, hash_id=0, seed_partition=0x9E3779B97F4A7C15, seed_table=1):
        R, S = np.ascontiguousarray(R), np.ascontiguousarray(S)
        res = OResult()
        rc = self.lib.phjo_join_radix(R.ctypes.data, R.shape[0], S.ctypes.data, S.shape[0], P, workers,
                                      table_kind, hash_id, u64(seed_partition), u64(seed_table), C.byref(res))
        assert rc == 0, rc
        return int(res.matches)

    def radix_partition(self, rel, P, hash_id, seed, workers=1, batch=None):
        rel = np.ascontiguousarray(rel)
        n = rel.shape[0]
        if batch is None:
            batch = n // workers if workers else 0
        out = np.empty(n, dtype=TUPLE)
        bounds = np.empty(2 * P, dtype=np.uint64)
        self.lib.phjo_radix_partition(rel.ctypes.data, n, P, hash_id, u64(seed & (2**64 - 1)), workers, batch,
                                      out.ctypes.data, bounds.ctypes.data, None)
        return out, bounds.reshape(P, 2)

    def partitioning_configuration(self, sizeA, sizeB, pool_workers, min_batch=10000):
        w, a, b = sz(), sz(), sz()
        self.lib.phjo_partitioning_configuration(sizeA, sizeB, pool_workers, min_batch, C.byref(w), C.byref(a), C.byref(b))
        return int(w.value), int(a.value), int(b.value)

    def join_materialize(self, R, S, table_kind=0, hash_id=0, seed_table=1):
        """All (probe tuple, equal-key build tuple) rows as a JOINED array (probe order)."""
        R, S = np.ascontiguousarray(R), np.ascontiguousarray(S)
        args = (R.ctypes.data, R.shape[0], S.ctypes.data, S.shape[0], table_kind, hash_id, u64(seed_table))
        n = int(self.lib.phjo_join_materialize(*args, None, 0))
        assert n != 2**64 - 1
        out = np.empty(n, dtype=JOINED)
        assert int(self.lib.phjo_join_materialize(*args, out.ctypes.data, n)) == n
        return out

    def count_by_sort(self, R, S):
        R, S = np.ascontiguousarray(R), np.ascontiguousarray(S)
        return int(self.lib.phjo_count_by_sort(R.ctypes.data, R.shape[0], S.ctypes.data, S.shape[0]))

    def table_probe(self, kind, size_ratio, hash_id, seed, number_of_objects, build, probe_keys):
        """Insert build tuples in order, then (exists, get index, getall count) per probe key."""
        build = np.ascontiguousarray(build)
        t = self.lib.phjo_table_new(kind, size_ratio, hash_id, u64(seed), number_of_objects)
        if not t:
            raise ValueError("numberOfObjects must be greater than zero")
        try:
            for i in range(build.shape[0]):
                rc = self.lib.phjo_table_insert(t, int(build["id"][i]), build.ctypes.data + 16 * i)
                if rc:
                    raise RuntimeError("BucketAllocator exceeded its limit.")
            buckets = int(self.lib.phjo_table_buckets(t))
            ex, gi, ga = [], [], []
            for k in probe_keys:
                ex.append(int(self.lib.phjo_table_exists(t, int(k))))
                g = self.lib.phjo_table_get(t, int(k))
                gi.append(-1 if not g else (g - build.ctypes.data) // 16)
                ga.append(int(self.lib.phjo_table_get_all(t, int(k), None, 0)))
            return buckets, ex, gi, ga
        finally:
            self.lib.phjo_table_free(t)


class Reference:
    """The unmodified reference behind oracle/ref_harness.cpp."""

    def __init__(self):
        path = build_ref()
        if not path or not os.path.exists(path):
            raise FileNotFoundError("oracle/_ref/libphj_ref.so is absent and /root/reference is not mounted")
        self.lib = lib = C.CDLL(path)
        _sig(lib, "phjref_last_error", C.c_char_p, [])
        _sig(lib, "phjref_default_workers", C.c_int, [])
        _sig(lib, "phjref_xxh3_64", u64, [i64, u64])
        _sig(lib, "phjref_hash", u64, [i64, u64, u64])
        _sig(lib, "phjref_lcg_doubles", None, [C.c_long, sz, vp])
        _sig(lib, "phjref_zipf_samples", C.c_int, [dbl, u64, C.c_long, sz, vp])
        _sig(lib, "phjref_table_new", vp, [sz])
        _sig(lib, "phjref_table_free", None, [vp])
        _sig(lib, "phjref_table_data", vp, [vp])
        _sig(lib, "phjref_fill_sequential", C.c_int, [vp, i64, C.c_int])
        _sig(lib, "phjref_fill_zipf_seeded", C.c_int, [vp, dbl, i64, i64, C.c_long, sz, C.c_int])
        _sig(lib, "phjref_join_tables", C.c_int, [vp, vp, C.c_int, C.c_int, sz, C.c_int, C.c_int, u64, u64, C.POINTER(RefResult)])
        _sig(lib, "phjref_join", C.c_int, [vp, sz, vp, sz, C.c_int, C.c_int, sz, C.c_int, C.c_int, u64, u64, C.POINTER(RefResult)])
        _sig(lib, "phjref_join_tables_forked", C.c_int, [vp, vp, C.c_int, C.c_int, sz, C.c_int, C.c_int, u64, u64, C.POINTER(RefResult)])
        _sig(lib, "phjref_table_probe", C.c_int, [C.c_int, dbl, u64, sz, vp, sz, C.c_int, vp, sz, vp, vp, vp])
        _sig(lib, "phjref_number_of_buckets", u64, [C.c_int, dbl, u64])
        _sig(lib, "phjref_format_json", C.c_int, [vp, vp, sz, u64, u64, u64, C.c_char_p, C.c_char_p, sz])
        _sig(lib, "phjref_join_type_from_string", C.c_int, [C.c_char_p])

    def default_workers(self):
        return int(self.lib.phjref_default_workers())

    def xxh3(self, key, seed):
        return int(self.lib.phjref_xxh3_64(int(key), u64(int(seed) & (2**64 - 1))))

    def lcg_doubles(self, seed, n):
        out = np.empty(n, dtype=np.float64)
        self.lib.phjref_lcg_doubles(seed, n, out.ctypes.data)
        return out.tolist()

    def zipf_samples(self, alpha, cardinality, seed, n):
        out = np.empty(n, dtype=np.uint64)
        if self.lib.phjref_zipf_samples(alpha, cardinality, seed, n, out.ctypes.data):
            raise ValueError(self.lib.phjref_last_error().decode())
        return out.tolist()

    def _table_array(self, h, n):
        ptr = self.lib.phjref_table_data(h)
        if not ptr or n == 0:
            return np.empty(0, dtype=TUPLE)
        buf = (C.c_char * (n * 16)).from_address(ptr)
        return np.frombuffer(buf, dtype=TUPLE, count=n)

    def fill_sequential(self, n, start=1, threads=3):
        h = self.lib.phjref_table_new(n)
        try:
            assert self.lib.phjref_fill_sequential(h, start, threads) == 0
            return self._table_array(h, n).copy()
        finally:
            self.lib.phjref_table_free(h)

    def fill_zipf(self, n, alpha, lo, hi, base_seed, batches, threads=3):
        h = self.lib.phjref_table_new(n)
        try:
            if self.lib.phjref_fill_zipf_seeded(h, alpha, lo, hi, base_seed, batches, threads):
                raise ValueError(self.lib.phjref_last_error().decode())
            return self._table_array(h, n).copy()
        finally:
            self.lib.phjref_table_free(h)

    def join(self, R, S, algo, partitions=32, table_kind=0, threads=0, seeded=True,
             seed_partition=0x9E3779B97F4A7C15, seed_table=1):
        """algo: 0 NoPartitioning::HashJoiner, 1 RadixClustering::HashJoiner. Returns a dict."""
        R, S = np.ascontiguousarray(R), np.ascontiguousarray(S)
        res = RefResult()
        rc = self.lib.phjref_join(R.ctypes.data, R.shape[0], S.ctypes.data, S.shape[0], algo, table_kind,
                                  partitions, threads, 1 if seeded else 0, u64(seed_partition), u64(seed_table),
                                  C.byref(res))
        if rc:
            raise RuntimeError(self.lib.phjref_last_error().decode())
        return {n: int(getattr(res, n)) for n, _ in res._fields_}

    # -- reference-owned tables: generated in place by the reference's generators and joined any number of
    # -- times without copies (bench.py --impl reference / cpu_baseline)
    def new_table(self, n):
        return self.lib.phjref_table_new(n)

    def free_table(self, h):
        self.lib.phjref_table_free(h)

    def table_view(self, h, n):
        """numpy view (no copy) of a reference table's tuples; valid until free_table."""
        return self._table_array(h, n)

    def generate_tables(self, n_build, n_probe, alpha, base_seed, batches, threads=0):
        """Sequential build side (ids 1..n) + seeded Zipf probe side over [1, n_build], as src/main.cpp:54-63."""
        hR, hS = self.new_table(n_build), self.new_table(n_probe)
        assert self.lib.phjref_fill_sequential(hR, 1, threads) == 0
        if self.lib.phjref_fill_zipf_seeded(hS, alpha, 1, n_build, base_seed, batches, threads):
            raise ValueError(self.lib.phjref_last_error().decode())
        return hR, hS

    def join_tables(self, hR, hS, algo, partitions=32, table_kind=0, threads=0, seeded=False,
                    seed_partition=0x9E3779B97F4A7C15, seed_table=1, forked=True):
        """Join two reference tables. forked: in a child process, because the reference's radix joiner keeps
        its partitioned tables alive after Run (3.36 GB per 10 M x 200 M join; see ref_harness.cpp)."""
        res = RefResult()
        fn = self.lib.phjref_join_tables_forked if forked else self.lib.phjref_join_tables
        rc = fn(hR, hS, algo, table_kind, partitions, threads, 1 if seeded else 0, u64(seed_partition),
                u64(seed_table), C.byref(res))
        if rc:
            raise RuntimeError(self.lib.phjref_last_error().decode())
        return {n: int(getattr(res, n)) for n, _ in res._fields_}

    def table_probe(self, kind, size_ratio, seed, number_of_objects, build, probe_keys, insert_threads=1):
        build = np.ascontiguousarray(build)
        probe = np.ascontiguousarray(probe_keys, dtype=np.int64)
        ex = np.zeros(probe.shape[0], dtype=np.uint8)
        gi = np.zeros(probe.shape[0], dtype=np.int64)
        ga = np.zeros(probe.shape[0], dtype=np.uint64)
        rc = self.lib.phjref_table_probe(kind, size_ratio, u64(seed), number_of_objects, build.ctypes.data,
                                         build.shape[0], insert_threads, probe.ctypes.data, probe.shape[0],
                                         ex.ctypes.data, gi.ctypes.data, ga.ctypes.data)
        if rc:
            raise RuntimeError(self.lib.phjref_last_error().decode())
        buckets = int(self.lib.phjref_number_of_buckets(kind, size_ratio, number_of_objects))
        return buckets, ex.tolist(), gi.tolist(), ga.tolist()

    def format_json(self, params: dict, partition_ns, build_ns, probe_ns, unit="ms"):
        keys = (C.c_char_p * len(params))(*[k.encode() for k in params])
        vals = (C.c_char_p * len(params))(*[str(v).encode() for v in params.values()])
        buf = C.create_string_buffer(4096)
        rc = self.lib.phjref_format_json(keys, vals, len(params), partition_ns, build_ns, probe_ns, unit.encode(), buf, 4096)
        if rc:
            raise RuntimeError(self.lib.phjref_last_error().decode())
        return buf.value.decode()
