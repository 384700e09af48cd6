"""Host logic of the staged host-pointer path (csrc/host_stage.cpp), checked without a GPU through
the test hook pumitally_debug_stage: what lands in the pinned slots, which origins are reported
as changed, that the caller's flying[] is zeroed (PumiTallyImpl.cpp:169-172), for both code paths
(AVX2 on aligned slots, scalar otherwise) and any number of workers."""
import ctypes as C

import numpy as np
import pytest

from pumiumtally_b200.tally import load_library

PATCH = np.dtype([("x", "f8"), ("y", "f8"), ("z", "f8"), ("idx", "i4"), ("pad", "i4")])


def aligned(n, dtype, offset_bytes=0):
    """n items whose address is 64-byte aligned plus offset_bytes."""
    item = np.dtype(dtype).itemsize
    raw = np.zeros(n * item + 128, dtype=np.uint8)
    start = (-raw.ctypes.data) % 64 + offset_bytes
    return raw[start:start + n * item].view(dtype)


def stage(origin, dest, flying, w, b_dest, b_w, b_fly, compare, threads, cap):
    L = load_library()
    out = np.zeros(max(cap, 1), dtype=PATCH)
    dp, bp = C.POINTER(C.c_double), C.POINTER(C.c_int8)
    n = len(w)
    rc = L.pumitally_debug_stage(origin.ctypes.data_as(dp), dest.ctypes.data_as(dp), flying.ctypes.data_as(bp),
                                 w.ctypes.data_as(dp), b_dest.ctypes.data_as(dp), b_w.ctypes.data_as(dp),
                                 b_fly.ctypes.data_as(bp), n, int(compare), threads, out.ctypes.data, cap)
    return rc, out


@pytest.mark.parametrize("threads", [1, 2, 7])
@pytest.mark.parametrize("n", [0, 1, 3, 4, 37, 4096, 100_003])
@pytest.mark.parametrize("misalign", [0, 8])
def test_stage_pass_matches_its_specification(threads, n, misalign):
    rng = np.random.default_rng(n + threads)
    origin, dest, w = rng.normal(size=3 * n), rng.normal(size=3 * n), rng.uniform(0.5, 1, n)
    flying = (rng.random(n) < 0.9).astype(np.int8)
    odd = rng.random(n) < 0.05
    flying[odd] = rng.choice(np.array([2, -1, 127, -128], dtype=np.int8), int(odd.sum()))
    b_dest, b_w, b_fly = aligned(3 * n, "f8", misalign), aligned(n, "f8", misalign), aligned(n, "i1")
    mirror = rng.normal(size=3 * n)
    same = rng.random(n) < 0.8                      # most origins equal the previous destination
    origin.reshape(-1, 3)[same] = mirror.reshape(-1, 3)[same]
    # differences in a single coordinate, in the sign of zero, NaNs
    if n > 8:
        origin[3 * 5 + 2] = np.nextafter(mirror[3 * 5 + 2], 9.0)
        mirror[3 * 6], origin[3 * 6] = 0.0, -0.0
        origin[3 * 7 + 1] = np.nan
        flying[5:8] = 1
    b_dest[:] = mirror
    old_w, old_fly = rng.normal(size=n), rng.integers(-3, 3, n).astype(np.int8)
    b_w[:], b_fly[:] = old_w, old_fly
    f_in = flying.copy()
    rc, out = stage(origin, dest, flying, w, b_dest, b_w, b_fly, True, threads, n + 1)
    fly = f_in == 1
    changed = fly & (origin.view(np.uint64).reshape(-1, 3) != mirror.view(np.uint64).reshape(-1, 3)).any(1)
    assert rc == int(changed.sum())
    got = np.sort(out[:rc], order="idx")
    np.testing.assert_array_equal(got["idx"], np.flatnonzero(changed))
    want = origin.reshape(-1, 3)[changed]
    for k, name in enumerate("xyz"):
        np.testing.assert_array_equal(got[name].view(np.uint64), want[:, k].copy().view(np.uint64))
    assert not flying.any()                                         # caller's flags zeroed
    np.testing.assert_array_equal(b_fly, f_in)                      # and preserved in the slots, as given
    # every slot is refilled, flying or not (branch-free copy; see host_stage.hpp)
    np.testing.assert_array_equal(b_dest.view(np.uint64), dest.view(np.uint64))
    np.testing.assert_array_equal(b_w, w)


def test_stage_pass_without_compare_and_with_overflow():
    n = 10_000
    rng = np.random.default_rng(1)
    origin, dest, w = rng.normal(size=3 * n), rng.normal(size=3 * n), rng.uniform(0.5, 1, n)
    b_dest, b_w, b_fly = aligned(3 * n, "f8"), aligned(n, "f8"), aligned(n, "i1")
    flying = np.ones(n, dtype=np.int8)
    rc, _ = stage(origin, dest, flying, w, b_dest, b_w, b_fly, False, 3, 16)
    assert rc == 0 and not flying.any()
    np.testing.assert_array_equal(b_dest, dest)
    # every origin differs from the mirror, room for 100: reported as overflow, slots refilled all the same
    flying[:] = 1
    dest2 = rng.normal(size=3 * n)
    rc, _ = stage(origin, dest2, flying, w, b_dest, b_w, b_fly, True, 3, 100)
    assert rc == -1 and not flying.any()
    np.testing.assert_array_equal(b_dest, dest2)
    np.testing.assert_array_equal(b_w, w)
