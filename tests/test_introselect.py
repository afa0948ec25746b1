"""The GPU's replay of libstdc++ std::nth_element (vs_introselect.cuh, compiled for the CPU
by tests/native/hooks.cpp) against the real std::nth_element / std::__introselect called by
the oracle.  The selected SUBSET AND ORDER must match (alignment.cpp:460-486, 526-545)."""
import ctypes as C

import numpy as np
import pytest


def pack(hooks, v):
    """keys = abs_delta << KEY_SHIFT | index (KEY_SHIFT = 17: abs_delta below 32768, up to 131071 tiles)"""
    shift = hooks.th_key_shift()
    assert int(v.max(initial=0)) < (1 << (32 - shift))
    return ((v.astype(np.uint32) << shift) | np.arange(len(v), dtype=np.uint32)).astype(np.uint32), (1 << shift) - 1


def replay(hooks, v, k):
    keys, mask = pack(hooks, v)
    hooks.th_nth_element(keys.ctypes.data_as(C.c_void_p), len(v), k)
    return keys & mask


@pytest.mark.parametrize("n", [1, 2, 3, 4, 5, 7, 16, 176, 220, 480, 1296, 2304, 5184, 20736, 82944])
def test_matches_std_nth_element(ob, hooks, n):
    rng = np.random.default_rng(n)
    for trial in range(12):
        kind = trial % 6
        if kind == 0:
            v = rng.integers(0, 256, n)
        elif kind == 1:   # residual-like: geometric, heavy ties at small values
            v = np.minimum(rng.exponential(6, n).astype(np.int64), 32767)
        elif kind == 2:
            v = rng.integers(0, 3, n)
        elif kind == 3:
            v = np.sort(rng.integers(0, 50, n))[:: (1 if trial % 2 else -1)]
        elif kind == 4:
            v = np.zeros(n, np.int64)
        else:
            v = rng.integers(0, 32768, n)
        v = v.astype(np.uint16)
        k = hooks.th_selected_count(n, C.c_float(0.8))
        ref = ob.select_smallest(v, 0.8)
        assert len(ref) == k
        assert np.array_equal(replay(hooks, v, k)[:k], ref)


def test_heap_select_fallback_matches_libstdcxx(ob, hooks):
    """Depth limit exhausted -> std::__heap_select; forced by calling __introselect with a
    small depth on both sides."""
    rng = np.random.default_rng(1)
    for trial in range(400):
        n = int(rng.integers(5, 3000))
        depth = int(rng.integers(0, 4))
        v = rng.integers(0, [4, 256, 32768][trial % 3], n).astype(np.uint16)
        nth = int(rng.integers(0, n))
        ref = ob.introselect_depth(v, nth, depth)
        keys, mask = pack(hooks, v)
        hooks.th_introselect_depth(keys.ctypes.data_as(C.c_void_p), n, nth, depth)
        assert np.array_equal(keys & mask, ref)


def test_selected_count_is_float_product(hooks):
    # (size_t)(N * 0.8f): float multiply, truncation (alignment.cpp:464-465)
    for n, k in ((5184, 4147), (20736, 16588), (1296, 1036), (1980, 1584), (480, 384), (5, 4), (1, 0)):
        assert hooks.th_selected_count(n, C.c_float(0.8)) == k
