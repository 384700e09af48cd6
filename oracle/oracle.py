"""ctypes wrapper over oracle/umtally_oracle.c.

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs.  Never imported by the
pumiumtally_b200 package.

The class mirrors the reference interface for the path (method names and
argument meaning of pumitally::PumiTally, reference: src/pumitally/PumiTally.h:50-103)
so parity tests read the same on both sides.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libumtally_oracle.so")
_lib = None


def build(force: bool = False) -> str:
    """Compile the C restatement (gcc, seconds)."""
    src = os.path.join(_HERE, "umtally_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        dp, ip, bp = C.POINTER(C.c_double), C.POINTER(C.c_int), C.POINTER(C.c_byte)
        L.um_oracle_create.restype = C.c_void_p
        L.um_oracle_create.argtypes = [dp, C.c_int, ip, C.c_int, C.c_int]
        L.um_oracle_destroy.argtypes = [C.c_void_p]
        L.um_oracle_set_mode.argtypes = [C.c_void_p, C.c_int]
        L.um_oracle_set_exit_rule.argtypes = [C.c_void_p, C.c_int]
        L.um_oracle_set_state.argtypes = [C.c_void_p, dp, ip, C.c_int]
        L.um_oracle_copy_initial_position.argtypes = [C.c_void_p, dp, C.c_int]
        L.um_oracle_move_to_next_location.argtypes = [C.c_void_p, dp, dp, bp, dp, C.c_int]
        L.um_oracle_normalized_flux.argtypes = [C.c_void_p, dp, dp]
        for name, rt in [
            ("um_oracle_flux", dp), ("um_oracle_elem_ids", ip), ("um_oracle_positions", dp),
            ("um_oracle_adjacency", ip),
        ]:
            getattr(L, name).restype = rt
            getattr(L, name).argtypes = [C.c_void_p]
        for name in ["um_oracle_n_segments", "um_oracle_n_crossings", "um_oracle_n_tracks", "um_oracle_n_lost"]:
            getattr(L, name).restype = C.c_longlong
            getattr(L, name).argtypes = [C.c_void_p]
        L.um_oracle_reset_flux.argtypes = [C.c_void_p]
        L.um_oracle_num_threads.restype = C.c_int
        L.um_oracle_set_num_threads.argtypes = [C.c_int]
        L.um_bruteforce_tally.argtypes = [dp, ip, C.c_int, dp, dp, dp, C.c_int, dp, dp, ip]
        _lib = L
    return _lib


def _d(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _i(a):
    return a.ctypes.data_as(C.POINTER(C.c_int))


class OraclePumiTally:
    """CPU oracle with the reference's method names.

    ``per_particle=False`` runs the reference-shaped global iterate-until-all-done
    loop; ``True`` walks each particle to completion (same arithmetic, faster).
    ``strict_exit`` selects the exit-face rule without the parallel-face tolerance.
    """

    def __init__(self, coords, tet2vert, num_particles: int, per_particle: bool = True, strict_exit: bool = False):
        self._L = lib()
        self.coords = np.ascontiguousarray(coords, dtype=np.float64)
        self.tet2vert = np.ascontiguousarray(tet2vert, dtype=np.int32)
        self.num_particles = int(num_particles)
        self.ntets = int(self.tet2vert.shape[0])
        self._h = self._L.um_oracle_create(
            _d(self.coords), self.coords.shape[0], _i(self.tet2vert), self.ntets, self.num_particles
        )
        self._L.um_oracle_set_mode(self._h, 1 if per_particle else 0)
        # strict_exit=True: the exit rule before the parallel-face tolerance (every face with n.u > 0 is a
        # candidate); identical results on generic inputs, kept to show exactly that
        self._L.um_oracle_set_exit_rule(self._h, 1 if strict_exit else 0)

    def __del__(self):
        if getattr(self, "_h", None):
            self._L.um_oracle_destroy(self._h)
            self._h = None

    # -- reference interface (PumiTally.h:66-95) -------------------------------
    def CopyInitialPosition(self, init_particle_positions, size=None):
        xyz = np.ascontiguousarray(init_particle_positions, dtype=np.float64).reshape(-1)
        size = xyz.size if size is None else size
        self._L.um_oracle_copy_initial_position(self._h, _d(xyz), int(size))

    def MoveToNextLocation(self, particle_origin, particle_destinations, flying, weights, size=None):
        o = np.ascontiguousarray(particle_origin, dtype=np.float64).reshape(-1)
        d = np.ascontiguousarray(particle_destinations, dtype=np.float64).reshape(-1)
        w = np.ascontiguousarray(weights, dtype=np.float64).reshape(-1)
        assert flying.dtype == np.int8 and flying.flags.c_contiguous, "flying must be int8 (it is written)"
        size = o.size if size is None else size
        self._L.um_oracle_move_to_next_location(
            self._h, _d(o), _d(d), flying.ctypes.data_as(C.POINTER(C.c_byte)), _d(w), int(size)
        )

    def set_state(self, xyz, elems):
        """Particles [0, len(elems)) placed at xyz in tets elems without walking (driver tests)."""
        xyz = np.ascontiguousarray(xyz, dtype=np.float64).reshape(-1)
        elems = np.ascontiguousarray(elems, dtype=np.int32)
        self._L.um_oracle_set_state(self._h, _d(xyz), _i(elems), len(elems))

    def normalized_flux(self):
        f = np.empty(self.ntets)
        v = np.empty(self.ntets)
        self._L.um_oracle_normalized_flux(self._h, _d(f), _d(v))
        return f, v

    # -- accessors (the reference tests reach into Impl members instead) --------
    @property
    def flux(self):
        return np.ctypeslib.as_array(self._L.um_oracle_flux(self._h), shape=(self.ntets,)).copy()

    @property
    def elem_ids(self):
        return np.ctypeslib.as_array(self._L.um_oracle_elem_ids(self._h), shape=(self.num_particles,)).copy()

    @property
    def positions(self):
        return np.ctypeslib.as_array(self._L.um_oracle_positions(self._h), shape=(self.num_particles, 3)).copy()

    @property
    def adjacency(self):
        return np.ctypeslib.as_array(self._L.um_oracle_adjacency(self._h), shape=(self.ntets, 4)).copy()

    @property
    def n_segments(self):
        return int(self._L.um_oracle_n_segments(self._h))

    @property
    def n_crossings(self):
        return int(self._L.um_oracle_n_crossings(self._h))

    @property
    def n_tracks(self):
        return int(self._L.um_oracle_n_tracks(self._h))

    @property
    def n_lost(self):
        return int(self._L.um_oracle_n_lost(self._h))

    def reset_flux(self):
        self._L.um_oracle_reset_flux(self._h)


def num_threads() -> int:
    return int(lib().um_oracle_num_threads())


def set_num_threads(n: int) -> None:
    lib().um_oracle_set_num_threads(int(n))


def bruteforce_tally(coords, tet2vert, a, b, w):
    """Independent O(N*E) integrator: returns (flux[E], t_last[N], elem_last[N])."""
    L = lib()
    coords = np.ascontiguousarray(coords, dtype=np.float64)
    t2v = np.ascontiguousarray(tet2vert, dtype=np.int32)
    a = np.ascontiguousarray(a, dtype=np.float64).reshape(-1, 3)
    b = np.ascontiguousarray(b, dtype=np.float64).reshape(-1, 3)
    w = np.ascontiguousarray(w, dtype=np.float64).reshape(-1)
    n = a.shape[0]
    flux = np.zeros(t2v.shape[0])
    tl = np.zeros(n)
    el = np.zeros(n, dtype=np.int32)
    L.um_bruteforce_tally(_d(coords), _i(t2v), t2v.shape[0], _d(a), _d(b), _d(w), n, _d(flux), _d(tl), _i(el))
    return flux, tl, el
