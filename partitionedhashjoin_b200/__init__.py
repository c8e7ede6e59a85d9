"""partitionedhashjoin_b200 -- B200-native hash-join engine behind phjoin's joiner interface.

Layout: ``csrc/`` CUDA kernels + the C ABI of ``include/phj.h`` (-> ``libphj_b200.so``);
``host/`` the C++ mirror of the reference's HashJoiner / CLI; ``engine.py`` the Python mirror used
by tests and bench. Importing this package requires the built shared library (no CPU fallback).
"""
from ._lib import (ALGO_NO_PARTITIONING, ALGO_RADIX_PARTITIONING, FLAG_CHAINED_TABLE, FLAG_COOP_PROBE, FLAG_COUNT_PIECEWISE, FLAG_COUNT_UPFRONT, FLAG_FUSE_HIST2, FLAG_HOT_DIGITS, FLAG_L2_TABLES,
                   FLAG_NO_FUSE_HIST2, FLAG_NO_HOT_DIGITS, FLAG_NO_HIST12, FLAG_NO_TMA_STORE, HASH_CITY,
                   HASH_MURMUR3, HASH_NAMES, HASH_XXH3, JOINED_DTYPE, LIB_PATH, TUPLE_DTYPE, PhjError)
from .engine import (DeviceTuples, Engine, Hasher, HashJoinTimer, HashJoinTimingResult, NoOpHashJoinTimer,
                     NoPartitioningConfiguration, NoPartitioningHashJoiner, PinnedTuples,
                     RadixClusteringConfiguration, RadixClusteringHashJoiner, as_tuples,
                     device_count, device_info, fill_sequential, fill_zipf, hash_batch, hash_host,
                     make_tuples)

__all__ = [n for n in dir() if not n.startswith("_")]
