import numpy as np
import pytest

from pumiumtally_b200.workload import CONFIGS, SyntheticWorkload


def _run(backend, n=5000, steps=3, **kw):
    wl = SyntheticWorkload(box=(6.0, 5.0, 4.0), num_particles=n, backend=backend, **kw)
    out = [np.asarray(wl.initial_positions())]
    for _ in range(steps):
        out.extend(np.asarray(a) for a in wl.next_step())
    return out


def test_numpy_and_torch_streams_are_identical():
    a, b = _run("numpy"), _run("torch")
    for x, y in zip(a, b):
        if x.dtype == np.float64:
            np.testing.assert_allclose(x, y, rtol=0, atol=1e-12)  # libm vs torch sin/cos/log: few ulp
        else:
            np.testing.assert_array_equal(x, y)


def test_subrange_generation_matches_full_range():
    full = SyntheticWorkload(box=(3.0, 3.0, 3.0), num_particles=1000)
    part = SyntheticWorkload(box=(3.0, 3.0, 3.0), num_particles=300, id_offset=500)
    np.testing.assert_array_equal(full.initial_positions()[500:800], part.initial_positions())
    for _ in range(2):
        f, p = full.next_step(), part.next_step()
        for x, y in zip(f, p):
            np.testing.assert_array_equal(x[500:800], y)


def test_workload_properties():
    wl = SyntheticWorkload(box=(6.0, 5.0, 4.0), num_particles=20000, mean_length=3.0)
    init = wl.initial_positions()
    assert (init > 0).all() and (init < np.array([6.0, 5.0, 4.0])).all()
    prev_dest, prev_fly = None, None
    cur = init.copy()
    for s in range(4):
        o, d, f, w = wl.next_step()
        fly = f == 1
        assert 0.93 < fly.mean() < 0.97
        assert ((w >= 0.5) & (w <= 1.0)).all()
        np.testing.assert_array_equal(o[~fly], cur[~fly])
        np.testing.assert_array_equal(d[~fly], o[~fly])
        length = np.linalg.norm(d - o, axis=1)[fly]
        assert 2.7 < length.mean() < 3.3
        inside = ((d >= 0) & (d <= np.array([6.0, 5.0, 4.0]))).all(1)
        cont = fly & ~wl.resample if s == 0 else None
        cur = np.where(fly[:, None], d, cur)
    assert set(CONFIGS) == {"c1", "c2", "c3", "c4", "c5"}
    assert np.prod(CONFIGS["c2"]["cells"]) * 6 == 998_250


def test_forward_peaked_directions():
    wl = SyntheticWorkload(box=(32.0, 32.0, 163.0), num_particles=4000, mean_length=1000.0, mu_min=0.9)
    wl.initial_positions()
    o, d, f, w = wl.next_step()
    u = (d - o)[f == 1]
    mu = u[:, 2] / np.linalg.norm(u, axis=1)
    assert mu.min() >= 0.9 - 1e-12
