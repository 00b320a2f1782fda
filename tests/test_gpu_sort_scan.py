"""The library's own radix sort (one kernel per digit, decoupled look-back) and single-pass scan against numpy, through
two development entry points of libpe_b200.so (peb_debug_sort_pairs / peb_debug_exclusive_scan).  Both feed VoxelGrid
and the grid build, whose bit-exactness tests cover them end to end; these cases aim at the tile boundaries, at skewed
digits (every key in one bin: one look-back chain carries everything) and at stability."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    from pose_estimation_b200 import pcl
    from pose_estimation_b200.pcl import lib

    lib.peb_debug_sort_pairs.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_int]
    lib.peb_debug_exclusive_scan.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]
    ctx = pcl.Context(0)
    yield ctx, lib
    ctx.close()


def _keys(kind, n, bits, rng):
    if kind == "uniform":
        return rng.integers(0, 1 << bits, n, dtype=np.uint64).astype(np.uint32)
    if kind == "one_bin":      # all keys equal: a pure stability / look-back test
        return np.full(n, (1 << bits) - 1, np.uint32)
    if kind == "two_values":
        return np.where(rng.random(n) < 0.5, 3, (1 << bits) - 2).astype(np.uint32)
    if kind == "sorted_runs":  # what a voxel grid of an organized cloud looks like: long runs, few distinct high digits
        return np.sort(rng.integers(0, 1 << bits, n, dtype=np.uint64).astype(np.uint32))[::-1].copy()
    raise ValueError(kind)


@pytest.mark.parametrize("n", [1, 2, 255, 1024, 1025, 4095, 4096, 4097, 70001, (1 << 20) - 1, (1 << 20) + 4097, 2332800])
@pytest.mark.parametrize("kind,bits", [("uniform", 32), ("uniform", 25), ("uniform", 9), ("one_bin", 17), ("two_values", 24),
                                       ("sorted_runs", 23)])
def test_sort_pairs_is_a_stable_sort(dev, n, kind, bits):
    ctx, lib = dev
    rng = np.random.default_rng(n * 31 + bits)
    keys = _keys(kind, n, bits, rng)
    vals = np.arange(n, dtype=np.uint32)
    k, v = keys.copy(), vals.copy()
    ctx.check(lib.peb_debug_sort_pairs(ctx.handle, k.ctypes.data, v.ctypes.data, n, bits))
    order = np.argsort(keys, kind="stable")
    assert np.array_equal(k, keys[order])
    assert np.array_equal(v, vals[order])  # equal keys keep ascending original index


@pytest.mark.parametrize("n", [0, 1, 2047, 2048, 2049, 500000, 3000001])
def test_exclusive_scan(dev, n):
    ctx, lib = dev
    rng = np.random.default_rng(n + 5)
    x = rng.integers(0, 4, n, dtype=np.uint32)
    out = np.zeros(max(n, 1), np.uint32)
    total = np.zeros(1, np.uint32)
    ctx.check(lib.peb_debug_exclusive_scan(ctx.handle, x.ctypes.data, out.ctypes.data, n, total.ctypes.data))
    ref = np.concatenate([[0], np.cumsum(x, dtype=np.uint64)]).astype(np.uint32)
    assert np.array_equal(out[:n], ref[:n]) and int(total[0]) == int(ref[n])
