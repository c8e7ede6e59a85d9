"""ctypes binding of include/phj.h (libphj_b200.so).

The library is built in tree by ``partitionedhashjoin_b200/csrc/Makefile`` (``__graft_entry__.build``).
There is deliberately no fallback: if the shared object is missing, importing this module raises.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
# PHJ_LIB selects another build of the same library (kernel-shape variants, tools/tune_shapes.py)
LIB_PATH = os.environ.get("PHJ_LIB") or os.path.join(PKG_DIR, "libphj_b200.so")

#: numpy view of ``phj_tuple`` == reference ``Common::Tuple`` (src/Common/Table.hpp:20-25)
TUPLE_DTYPE = np.dtype([("id", "<i8"), ("payload", "<i8")], align=True)
assert TUPLE_DTYPE.itemsize == 16
#: ``phj_joined_tuple`` == reference ``Common::JoinedTuple`` (src/Common/Table.hpp:27-33)
JOINED_DTYPE = np.dtype([("id", "<i8"), ("payloadA", "<i8"), ("payloadB", "<i8")])
assert JOINED_DTYPE.itemsize == 24

ALGO_NO_PARTITIONING = 0
ALGO_RADIX_PARTITIONING = 1
ALGO_SHARD_SPLIT = 2
HASH_XXH3, HASH_MURMUR3, HASH_CITY = 0, 1, 2
HASH_NAMES = {"xxh3": HASH_XXH3, "xxhash": HASH_XXH3, "murmur3": HASH_MURMUR3, "city": HASH_CITY}
FLAG_NO_TMA_STORE = 0x2
FLAG_FUSE_HIST2 = 0x4
FLAG_SPLIT_REMOTE_ONLY = 0x8
FLAG_CHAINED_TABLE = 0x10
FLAG_NO_FUSE_HIST2 = 0x20
FLAG_SPLIT_LOCAL_TILES = 0x40
FLAG_L2_TABLES = 0x80
FLAG_NO_HIST12 = 0x100
FLAG_COOP_PROBE = 0x200
FLAG_COUNT_UPFRONT = 0x400
FLAG_COUNT_PIECEWISE = 0x800
FLAG_HOT_DIGITS = 0x1000
FLAG_NO_HOT_DIGITS = 0x2000

OK, ERR_INVALID, ERR_CUDA, ERR_STATE, ERR_NOMEM = 0, 1, 2, 3, 4


class PhjConfig(C.Structure):
    _fields_ = [
        ("algo", C.c_int32),
        ("hash", C.c_int32),
        ("partitions", C.c_uint64),
        ("radix_bits", C.c_uint32 * 2),
        ("hash_seed", C.c_uint64),
        ("table_seed", C.c_uint64),
        ("device", C.c_int32),
        ("flags", C.c_uint32),
        ("reserve_build", C.c_uint64),
        ("reserve_probe", C.c_uint64),
        ("shard_shift", C.c_uint32),
        ("split_ctas", C.c_uint32),
        ("split_chunks", C.c_uint32),
        ("upload_chunks", C.c_uint32),
        ("num_gpus", C.c_int32),
        ("reserved0", C.c_uint32),
    ]


class PhjResult(C.Structure):
    _fields_ = [
        ("matches", C.c_uint64),
        ("partition_ns", C.c_uint64),
        ("build_ns", C.c_uint64),
        ("probe_ns", C.c_uint64),
        ("join_ns", C.c_uint64),
        ("total_ns", C.c_uint64),
        ("h2d_ns", C.c_uint64),
        ("hbm_bytes_alg", C.c_uint64),
        ("kernel_launches", C.c_uint32),
        ("passes", C.c_uint32),
        ("partitions", C.c_uint64),
        ("fallback_partitions", C.c_uint64),
        ("h2d_bytes", C.c_uint64),
        ("d2h_bytes", C.c_uint64),
        ("joined_tuples", C.c_uint64),
        ("materialize_ns", C.c_uint64),
        ("e2e_ns", C.c_uint64),
        ("upload_chunks", C.c_uint32),
        ("gpus", C.c_uint32),
        ("count_ns", C.c_uint64),
        ("shuffle_ns", C.c_uint64),
        ("shuffle_bytes", C.c_uint64),
    ]

    def as_dict(self):
        return {name: int(getattr(self, name)) for name, _ in self._fields_}


class PhjDistLayout(C.Structure):
    _fields_ = [
        ("world", C.c_uint32),
        ("rank", C.c_uint32),
        ("digits", C.c_uint32),
        ("local_partitions", C.c_uint32),
        ("chunks", C.c_uint32),
        ("region_buckets", C.c_uint32),
        ("window_tuples", C.c_uint64 * 2),
        ("resizes", C.c_uint32),
        ("partitions_here", C.c_uint32),
        ("sent_remote_bytes", C.c_uint64),
        ("hot_count", C.c_uint32),
        ("hot_digits", C.c_uint32 * 32),
        ("reserved", C.c_uint32),
    ]


class PhjDeviceInfo(C.Structure):
    _fields_ = [
        ("name", C.c_char * 128),
        ("sm_count", C.c_int32),
        ("cc_major", C.c_int32),
        ("cc_minor", C.c_int32),
        ("global_mem_bytes", C.c_uint64),
        ("l2_bytes", C.c_uint64),
        ("smem_per_block_optin", C.c_uint64),
        ("sm_clock_khz", C.c_int32),
        ("mem_clock_khz", C.c_int32),
        ("mem_bus_bits", C.c_int32),
    ]


#: every symbol include/phj.h declares: name -> (restype, argtypes)
SIGNATURES = {
    "phj_create": (C.c_int, [C.POINTER(PhjConfig), C.POINTER(C.c_void_p)]),
    "phj_destroy": (None, [C.c_void_p]),
    "phj_last_error": (C.c_char_p, []),
    "phj_abi_version": (C.c_uint32, []),
    "phj_upload": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t]),
    "phj_bind_device": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t]),
    "phj_bind_device_partitioned": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t,
                                              C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32]),
    "phj_join": (C.c_int, [C.c_void_p, C.POINTER(PhjResult)]),
    "phj_join_materialize": (C.c_int, [C.c_void_p, C.POINTER(PhjResult)]),
    "phj_read_joined": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint64]),
    "phj_device_joined": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_uint64)]),
    "phj_join_host": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t,
                                C.POINTER(PhjResult)]),
    "phj_hash_batch": (C.c_int, [C.c_int32, C.c_uint64, C.c_void_p, C.c_size_t, C.c_void_p, C.c_int32]),
    "phj_hash_host": (C.c_uint64, [C.c_int32, C.c_uint64, C.c_int64]),
    "phj_read_partitions": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p]),
    "phj_device_partitions": (C.c_int, [C.c_void_p, C.c_int32, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p),
                                        C.POINTER(C.c_size_t)]),
    "phj_shard_count": (C.c_int, [C.c_void_p, C.c_void_p]),
    "phj_shard_scatter": (C.c_int, [C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                    C.POINTER(PhjResult)]),
    "phj_enable_peer_access": (C.c_int, [C.c_int32, C.c_int32]),
    "phj_nccl_unique_id": (C.c_int, [C.c_void_p]),
    "phj_dist_create": (C.c_int, [C.POINTER(PhjConfig), C.c_int32, C.c_int32, C.c_void_p, C.POINTER(C.c_void_p)]),
    "phj_dist_destroy": (None, [C.c_void_p]),
    "phj_dist_upload": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t]),
    "phj_dist_bind_device": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t]),
    "phj_dist_join": (C.c_int, [C.c_void_p, C.POINTER(PhjResult)]),
    "phj_dist_kernel_times": (C.c_int, [C.c_void_p, C.POINTER(C.c_char_p), C.POINTER(C.c_uint64), C.c_uint32]),
    "phj_dist_kernel_timing": (C.c_int, [C.c_void_p, C.c_char_p]),
    "phj_dist_kernel_trace": (C.c_int, [C.c_void_p, C.POINTER(C.c_char_p), C.POINTER(C.c_uint64), C.POINTER(C.c_uint64),
                                        C.c_uint32]),
    "phj_dist_measure_peer_copy": (C.c_int, [C.c_void_p, C.c_uint64, C.c_uint32, C.POINTER(C.c_uint64)]),
    "phj_dist_info": (C.c_int, [C.c_void_p, C.POINTER(PhjDistLayout)]),
    "phj_dist_read_window": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_uint64, C.c_void_p]),
    "phj_shared_alloc": (C.c_int, [C.c_int32, C.c_size_t, C.POINTER(C.c_void_p), C.c_void_p]),
    "phj_shared_open": (C.c_int, [C.c_int32, C.c_void_p, C.POINTER(C.c_void_p)]),
    "phj_shared_close": (C.c_int, [C.c_int32, C.c_void_p]),
    "phj_shared_free": (C.c_int, [C.c_int32, C.c_void_p]),
    "phj_device_fill_sequential": (C.c_int, [C.c_int32, C.c_void_p, C.c_size_t, C.c_int64]),
    "phj_device_fill_zipf": (C.c_int, [C.c_int32, C.c_void_p, C.c_size_t, C.c_double, C.c_int64, C.c_int64,
                                       C.c_int64, C.c_size_t]),
    "phj_memcpy_h2d": (C.c_int, [C.c_int32, C.c_void_p, C.c_void_p, C.c_size_t]),
    "phj_memcpy_d2h": (C.c_int, [C.c_int32, C.c_void_p, C.c_void_p, C.c_size_t]),
    "phj_memcpy_d2d": (C.c_int, [C.c_int32, C.c_void_p, C.c_void_p, C.c_size_t]),
    "phj_set_parent_digits": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint32]),
    "phj_kernel_timing": (C.c_int, [C.c_void_p, C.c_char_p]),
    "phj_kernel_times": (C.c_int, [C.c_void_p, C.POINTER(C.c_char_p), C.POINTER(C.c_uint64), C.c_uint32]),
    "phj_get_device_info": (C.c_int, [C.c_int32, C.POINTER(PhjDeviceInfo)]),
    "phj_device_count": (C.c_int, []),
    "phj_fill_sequential": (C.c_int, [C.c_void_p, C.c_size_t, C.c_int64, C.c_int32]),
    "phj_fill_zipf": (C.c_int, [C.c_void_p, C.c_size_t, C.c_double, C.c_int64, C.c_int64, C.c_int64,
                                C.c_size_t, C.c_int32]),
    "phj_host_alloc": (C.c_int, [C.POINTER(C.c_void_p), C.c_size_t]),
    "phj_host_free": (C.c_int, [C.c_void_p]),
}


class PhjError(RuntimeError):
    def __init__(self, status: int, message: str):
        super().__init__(f"phj error {status}: {message}")
        self.status = status
        self.message = message


def load_library(path: str = LIB_PATH) -> C.CDLL:
    if not os.path.exists(path):
        raise ImportError(
            f"{path} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "or `make -C partitionedhashjoin_b200/csrc`. There is no CPU fallback.")
    lib = C.CDLL(path)
    for name, (restype, argtypes) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is not exported
        fn.restype = restype
        fn.argtypes = argtypes
    return lib


lib = load_library()


def check(status: int) -> None:
    if status != OK:
        raise PhjError(status, lib.phj_last_error().decode("utf-8", "replace"))
