"""ctypes access to the CPU checkers under oracle/ -- TEST INFRASTRUCTURE ONLY.

  load_oracle(): oracle/librn_oracle.so  (C restatement, rn_oracle.c)
  load_ref():    oracle/_ref/libredref.so (the reference's own sources; None if not built)

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
import this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
ORACLE_SO = os.path.join(ORACLE_DIR, "librn_oracle.so")
REF_SO = os.path.join(ORACLE_DIR, "_ref", "libredref.so")
REF_TREE = "/root/reference/renderEngine"


class Params(C.Structure):
    _fields_ = [
        ("seed", C.c_uint32), ("depth", C.c_int32), ("amplitude", C.c_float), ("size", C.c_float),
        ("persistence", C.c_float), ("lacunarity", C.c_int32), ("sideways_scale", C.c_float),
        ("radiosity_strength", C.c_float), ("feature_scale", C.c_float), ("amplitude_variation", C.c_float),
        ("bulgy_rocks", C.c_int32), ("apply_volcano", C.c_int32),
    ]


class Cell(C.Structure):
    """rn_cell (include/rednoise.h): 32 bytes."""
    _fields_ = [("position", C.c_float * 3), ("colour", C.c_float * 3), ("radius", C.c_float), ("normal", C.c_int8 * 3),
                ("level", C.c_uint8)]


CELL_DTYPE = np.dtype([("position", np.float32, 3), ("colour", np.float32, 3), ("radius", np.float32), ("normal", np.int8, 3),
                       ("level", np.uint8)])


def cell_count(depth: int) -> int:
    return (4 ** depth - 1) // 3


def params(seed=12345, depth=8, apply_volcano=1, **kw) -> Params:
    d = dict(seed=seed, depth=depth, amplitude=0.2, size=500.0, persistence=0.5, lacunarity=2, sideways_scale=0.5,
             radiosity_strength=0.5, feature_scale=80.0, amplitude_variation=0.5, bulgy_rocks=1,
             apply_volcano=apply_volcano)
    d.update(kw)
    return Params(**d)


def build_oracle(force: bool = False):
    src = [os.path.join(ORACLE_DIR, f) for f in ("rn_oracle.c", "rn_oracle.h", "rn_oracle3d.c")]
    stale = force or not os.path.exists(ORACLE_SO) or any(
        os.path.exists(s) and os.path.getmtime(s) > os.path.getmtime(ORACLE_SO) for s in src)
    if stale:
        subprocess.run(["make", "-s", "-C", ORACLE_DIR, "oracle"], check=True)
    return ORACLE_SO


def build_ref():
    """Compiles the reference's own sources where they lie (only possible in the build container)."""
    if os.path.isdir(os.path.join(REF_TREE, "source")):
        subprocess.run(["make", "-s", "-C", ORACLE_DIR, "ref"], check=True)
    return REF_SO if os.path.exists(REF_SO) else None


class Oracle:
    def __init__(self, L):
        self.L = L
        vp = C.c_void_p
        L.rno_lcg_next_state.restype = C.c_uint32
        L.rno_lcg_next_state.argtypes = [C.c_uint32]
        L.rno_lcg_state_after.restype = C.c_uint32
        L.rno_lcg_state_after.argtypes = [C.c_uint32, C.c_uint64]
        L.rno_lcg_draw.restype = C.c_uint32
        L.rno_lcg_draw.argtypes = [C.c_uint32, C.c_uint64]
        L.rno_random_from_draw.restype = C.c_float
        L.rno_random_from_draw.argtypes = [C.c_int]
        L.rno_draw_count.restype = C.c_uint64
        L.rno_draw_count.argtypes = [C.c_int]
        L.rno_call_index.restype = C.c_uint64
        L.rno_call_index.argtypes = [C.c_int, C.c_int, C.c_int]
        L.rno_terrain2d_seq.argtypes = [vp, vp]
        L.rno_terrain2d_indexed.argtypes = [vp, vp, C.c_int]
        L.rno_add_volcano.restype = None
        L.rno_add_volcano.argtypes = [vp, vp, vp]
        L.rno_terrain2d.argtypes = [vp, vp, vp]
        L.rno_terrain2d_batch.argtypes = [vp, vp, C.c_int, vp, C.c_int]
        L.rno_cell_count.restype = C.c_uint64
        L.rno_cell_count.argtypes = [C.c_int]
        L.rno_terrain2d_cells.argtypes = [vp, vp, vp]
        L.rno_volume3d.argtypes = [C.c_uint32, C.c_int, C.c_float, C.c_int, C.c_int, vp]
        L.rno_rock_draw_count.restype = C.c_uint64
        L.rno_rock_draw_count.argtypes = [C.c_int]
        L.rno_rock.restype = C.c_longlong
        L.rno_rock.argtypes = [C.c_uint32, C.c_int, C.c_float, C.c_float, vp, vp, vp]
        L.rno_cloud.restype = C.c_longlong
        L.rno_cloud.argtypes = [C.c_uint32, C.c_int, C.c_float, vp, vp, vp]
        L.rno_rock_indexed.restype = C.c_longlong
        L.rno_rock_indexed.argtypes = [C.c_uint32, C.c_int, C.c_float, vp, vp]
        L.rno_fnv64.restype = C.c_uint64
        L.rno_fnv64.argtypes = [vp, C.c_size_t]

    def terrain2d(self, p: Params, want_volcano_pos=False):
        s = (1 << p.depth) + 1
        out = np.empty((s, s, 4), np.float32)
        vpos = (C.c_float * 3)()
        st = self.L.rno_terrain2d(C.byref(p), out.ctypes.data, vpos)
        if st != 0:
            raise RuntimeError(f"oracle status {st}")
        return (out, tuple(vpos)) if want_volcano_pos else out

    def terrain2d_indexed(self, p: Params, reverse=False):
        s = (1 << p.depth) + 1
        out = np.empty((s, s, 4), np.float32)
        st = self.L.rno_terrain2d_indexed(C.byref(p), out.ctypes.data, 1 if reverse else 0)
        if st != 0:
            raise RuntimeError(f"oracle status {st}")
        return out

    def terrain2d_batch(self, p: Params, seeds, threads=1):
        s = (1 << p.depth) + 1
        out = np.empty((len(seeds), s, s, 4), np.float32)
        arr = (C.c_uint32 * len(seeds))(*[int(x) & 0xFFFFFFFF for x in seeds])
        st = self.L.rno_terrain2d_batch(C.byref(p), arr, len(seeds), out.ctypes.data, threads)
        if st != 0:
            raise RuntimeError(f"oracle status {st}")
        return out

    def cells(self, p: Params, field: np.ndarray) -> np.ndarray:
        """Terrain::generateRecursive restated: the Cell quadtree of the finished field (structured array, CELL_DTYPE)."""
        field = np.ascontiguousarray(field, np.float32)
        out = np.zeros(cell_count(p.depth), CELL_DTYPE)
        st = self.L.rno_terrain2d_cells(C.byref(p), field.ctypes.data, out.ctypes.data)
        if st != 0:
            raise RuntimeError(f"oracle status {st}")
        return out

    def volume3d(self, seed: int, depth: int, amplitude: float, i0: int = 0, ni=None) -> np.ndarray:
        """3D extension (rn_oracle3d.c, no reference counterpart): planes [i0, i0+ni) of the (W+1)^3 volume, shape (ni, S, S)."""
        S = (1 << depth) + 1
        ni = S - i0 if ni is None else ni
        v = np.full((ni, S, S), np.nan, np.float32)
        st = self.L.rno_volume3d(seed & 0xFFFFFFFF, depth, amplitude, i0, ni, v.ctypes.data)
        if st != 0:
            raise RuntimeError(f"oracle status {st}")
        return v

    def rock(self, state: int, depth: int, amplitude: float = 0.3, want_cells: bool = True):
        """Rock.cpp restated: (faces[6, W+1, W+1, 2] = {height, shade} through Rock::element, cells[6, count], draws, state_after)."""
        S = (1 << depth) + 1
        faces = np.empty((6, S, S, 2), np.float32)
        cells = np.zeros((6, cell_count(depth)), CELL_DTYPE) if want_cells else None
        st = C.c_uint32(0)
        n = self.L.rno_rock(state & 0xFFFFFFFF, depth, amplitude, 1.0, faces.ctypes.data, cells.ctypes.data if want_cells else None, C.byref(st))
        if n < 0:
            raise RuntimeError(f"oracle status {n}")
        return faces, cells, int(n), int(st.value)

    def cloud(self, state: int, depth: int, amplitude: float = 0.3, want_cells: bool = True):
        """Cloud.cpp restated: (faces[2, W+1, W+1, 2], cells[2, count], draws, state_after)."""
        S = (1 << depth) + 1
        faces = np.empty((2, S, S, 2), np.float32)
        cells = np.zeros((2, cell_count(depth)), CELL_DTYPE) if want_cells else None
        st = C.c_uint32(0)
        n = self.L.rno_cloud(state & 0xFFFFFFFF, depth, amplitude, faces.ctypes.data, cells.ctypes.data if want_cells else None, C.byref(st))
        if n < 0:
            raise RuntimeError(f"oracle status {n}")
        return faces, cells, int(n), int(st.value)

    def rock_indexed(self, state: int, depth: int, amplitude: float = 0.3):
        """Closed-form call indices + breadth-first levels (the CUDA formulation): (faces, status, state_after)."""
        S = (1 << depth) + 1
        faces = np.empty((6, S, S, 2), np.float32)
        st = C.c_uint32(0)
        n = self.L.rno_rock_indexed(state & 0xFFFFFFFF, depth, amplitude, faces.ctypes.data, C.byref(st))
        return faces, int(n), int(st.value)

    def fnv64(self, a: np.ndarray) -> str:
        a = np.ascontiguousarray(a)
        return "%016x" % self.L.rno_fnv64(a.ctypes.data, a.nbytes)


class Ref:
    def __init__(self, L):
        self.L = L
        L.ref_terrain_passes.restype = C.c_longlong
        L.ref_terrain_passes.argtypes = [C.c_int, C.c_float, C.c_float, C.c_uint, C.c_int, C.c_void_p, C.c_void_p]
        L.ref_lcg_draws.restype = None
        L.ref_lcg_draws.argtypes = [C.c_uint, C.c_int, C.c_void_p]
        L.ref_terrain_cells.restype = C.c_longlong
        L.ref_terrain_cells.argtypes = [C.c_int, C.c_float, C.c_float, C.c_uint, C.c_void_p, C.c_void_p]
        L.ref_random_from_state.restype = C.c_float
        L.ref_random_from_state.argtypes = [C.c_uint]
        if hasattr(L, "ref_cloud"):
            L.ref_cloud.restype = C.c_longlong
            L.ref_cloud.argtypes = [C.c_uint, C.c_int, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p]
        if hasattr(L, "ref_rock"):
            L.ref_rock.restype = C.c_longlong
            L.ref_rock.argtypes = [C.c_uint, C.c_int, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p]

    def terrain2d(self, depth, seed=12345, amplitude=0.2, size=500.0, volcano=1, want_times=False):
        s = (1 << depth) + 1
        out = np.empty((s, s, 4), np.float32)
        times = (C.c_double * 3)()
        calls = self.L.ref_terrain_passes(depth, amplitude, size, seed, volcano, out.ctypes.data, times)
        if calls < 0:
            raise RuntimeError("reference harness rejected the arguments")
        return (out, calls, tuple(times)) if want_times else (out, calls)

    def cells(self, depth, seed=12345, amplitude=0.2, size=500.0):
        """The reference's own generateRecursive on its own field: (field, cells)."""
        s = (1 << depth) + 1
        field = np.empty((s, s, 4), np.float32)
        out = np.zeros(cell_count(depth), CELL_DTYPE)
        n = self.L.ref_terrain_cells(depth, amplitude, size, seed, field.ctypes.data, out.ctypes.data)
        if n != cell_count(depth):
            raise RuntimeError(f"reference harness returned {n}")
        return field, out

    def rock(self, state, depth, amplitude=0.3):
        """The reference's own Rock::generateBaseRecursive / generateRecursive: (faces, cells, draws, state_after)."""
        S = (1 << depth) + 1
        faces = np.empty((6, S, S, 2), np.float32)
        cells = np.zeros((6, cell_count(depth)), CELL_DTYPE)
        st = C.c_uint(0)
        n = self.L.ref_rock(state & 0xFFFFFFFF, depth, amplitude, faces.ctypes.data, cells.ctypes.data, C.byref(st))
        if n < 0:
            raise RuntimeError(f"reference harness returned {n}")
        return faces, cells, int(n), int(st.value)

    def cloud(self, state, depth, amplitude=0.3):
        S = (1 << depth) + 1
        faces = np.empty((2, S, S, 2), np.float32)
        cells = np.zeros((2, cell_count(depth)), CELL_DTYPE)
        st = C.c_uint(0)
        n = self.L.ref_cloud(state & 0xFFFFFFFF, depth, amplitude, faces.ctypes.data, cells.ctypes.data, C.byref(st))
        if n < 0:
            raise RuntimeError(f"reference harness returned {n}")
        return faces, cells, int(n), int(st.value)

    def lcg_draws(self, seed, n):
        out = np.empty(n, np.int32)
        self.L.ref_lcg_draws(seed, n, out.ctypes.data)
        return out


_oracle = None
_ref = None


def load_oracle() -> Oracle:
    global _oracle
    if _oracle is None:
        _oracle = Oracle(C.CDLL(build_oracle()))
    return _oracle


def load_ref():
    global _ref
    if _ref is None:
        if not os.path.exists(REF_SO):
            try:
                build_ref()
            except Exception:
                return None
        if not os.path.exists(REF_SO):
            return None
        _ref = Ref(C.CDLL(REF_SO))
    return _ref
