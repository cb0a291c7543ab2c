"""cpptraj_b200 -- B200-native best-fit RMSD path behind cpptraj's rms2d / cluster
pairwise cache / rmsd action.

The product is the CUDA library `libb200rmsd.so` (C ABI in include/b200_rmsd.h).
This Python package is a thin ctypes binding used by the tests and bench.py; it
never computes anything on the CPU and raises if the library is missing.
"""
from .api import (  # noqa: F401
    B200Error, lib, init, shutdown, num_devices, shard_rows, rms2d_tri, rms2d_tri_shard,
    rms2d_full, Rmsd1vN, rmsd_1vN, frames_to_centroids, build_centroids, hieragglo, rmsavgcorr, cache_resident_begin, cache_resident_end, cache_cluster_sums, cache_cluster_links, coords_resident_begin, coords_resident_end, set_profiling, reset_stats, get_stats, measure_fp64_mma_peak, set_mma_variant,
    dev_rms2d_tri, dev_rmsd_1vN, tri_size, tri_index, LIB_PATH, set_pair_engine, set_i8_cta_group, get_i8_cta_group, last_pair_engine, debug_i8, measure_i8_mma_peak, set_fixed_point_bits,
)
