"""ctypes binding of include/rednoise.h (the C ABI of librednoise_b200.so).

This is plumbing for the tests and bench.py, mirroring the header one to one.  The
product is the shared library; it has no CPU fallback, so loading fails loudly when
the library has not been built (python -m rednoise_b200.build).
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("RN_LIB") or os.path.join(HERE, "librednoise_b200.so")  # RN_LIB: build experiments only

RN_OK, RN_ERR_BAD_PARAM, RN_ERR_UNSUPPORTED, RN_ERR_NO_DEVICE, RN_ERR_CUDA, RN_ERR_OOM = range(6)
RN_MIN_DEPTH, RN_MAX_DEPTH = 3, 14

# every symbol include/rednoise.h declares (tests/test_abi_host.py::test_header_symbols_exported checks the two agree)
SYMBOLS = [
    "rn_version", "rn_last_error_string", "rn_status_string", "rn_params_default", "rn_terrain2d_side",
    "rn_terrain2d_draw_count", "rn_terrain2d_call_index", "rn_lcg_state_after", "rn_lcg_draw", "rn_create",
    "rn_destroy", "rn_launch_count", "rn_set_algorithm", "rn_set_profiling", "rn_last_stage_times", "rn_set_pipelining", "rn_terrain2d", "rn_terrain2d_batch", "rn_terrain2d_batch_chunk", "rn_lcg_draws_device", "rn_selftest_math",
    "rn_terrain2d_cell_count", "rn_terrain2d_cell_level_offset", "rn_terrain2d_cells", "rn_volume3d", "rn_volume3d_host", "rn_volume3d_slab_elems",
    "rn_rock_draw_count", "rn_rock", "rn_rock_batch", "rn_cloud_draw_count", "rn_cloud",
    "rn_device_count", "rn_device_alloc", "rn_device_free", "rn_host_alloc", "rn_host_free", "rn_copy_to_host",
    "rn_create_multi", "rn_destroy_multi", "rn_multi_size", "rn_multi_context", "rn_slab_rows", "rn_slab_window",
    "rn_terrain2d_sharded", "rn_terrain2d_batch_sharded", "rn_gather",
]


class Params(C.Structure):
    """rn_params: the reference's generator call surface (Terrain.h:20 + the literals of Terrain.cpp)."""

    _fields_ = [
        ("seed", C.c_uint32), ("depth", C.c_int32), ("amplitude", C.c_float), ("size", C.c_float),
        ("persistence", C.c_float), ("lacunarity", C.c_int32), ("sideways_scale", C.c_float),
        ("radiosity_strength", C.c_float), ("feature_scale", C.c_float), ("amplitude_variation", C.c_float),
        ("bulgy_rocks", C.c_int32), ("apply_volcano", C.c_int32),
    ]


CELL_DTYPE = np.dtype([("position", np.float32, 3), ("colour", np.float32, 3), ("radius", np.float32), ("normal", np.int8, 3),
                       ("level", np.uint8)])  # rn_cell, 32 bytes


class VolumeParams(C.Structure):
    """rn_volume_params (3D extension)."""
    _fields_ = [("seed", C.c_uint32), ("depth", C.c_int32), ("amplitude", C.c_float)]


class RockParams(C.Structure):
    """rn_rock_params: Rock(depth, amplitude, size) (Rock.h:10) + the CRT generator state at entry."""
    _fields_ = [("lcg_state", C.c_uint32), ("depth", C.c_int32), ("amplitude", C.c_float)]


ROCK_ELEMENT_DTYPE = np.dtype([("height", np.float32), ("shade", np.float32)])  # rn_rock_element (Rock.h:21-24)


class Window(C.Structure):
    _fields_ = [("i0", C.c_int32), ("j0", C.c_int32), ("ni", C.c_int32), ("nj", C.c_int32)]


class RedNoiseError(RuntimeError):
    def __init__(self, status: int, msg: str):
        super().__init__(f"rednoise status {status}: {msg}")
        self.status = status


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(f"{LIB_PATH} is missing: build it with `python -m rednoise_b200.build` "
                          "(there is no CPU fallback)")
    L = C.CDLL(LIB_PATH)
    vp, u32, u64, i32, sz = C.c_void_p, C.c_uint32, C.c_uint64, C.c_int, C.c_size_t
    L.rn_version.restype = i32
    L.rn_last_error_string.restype = C.c_char_p
    L.rn_status_string.restype = C.c_char_p
    L.rn_status_string.argtypes = [i32]
    L.rn_params_default.argtypes = [C.POINTER(Params)]
    L.rn_terrain2d_side.argtypes = [i32]
    L.rn_terrain2d_draw_count.restype = u64
    L.rn_terrain2d_draw_count.argtypes = [i32]
    L.rn_terrain2d_call_index.restype = u64
    L.rn_terrain2d_call_index.argtypes = [i32, i32, i32]
    L.rn_lcg_state_after.restype = u32
    L.rn_lcg_state_after.argtypes = [u32, u64]
    L.rn_lcg_draw.restype = u32
    L.rn_lcg_draw.argtypes = [u32, u64]
    L.rn_create.argtypes = [i32, C.POINTER(vp)]
    L.rn_destroy.argtypes = [vp]
    L.rn_launch_count.restype = u64
    L.rn_launch_count.argtypes = [vp]
    L.rn_set_algorithm.argtypes = [vp, i32]
    L.rn_set_profiling.argtypes = [vp, i32]
    L.rn_set_pipelining.argtypes = [vp, i32]
    L.rn_last_stage_times.argtypes = [vp, C.POINTER(C.c_float), i32]
    L.rn_terrain2d.argtypes = [vp, C.POINTER(Params), C.POINTER(Window), vp, sz, i32, vp, C.POINTER(C.c_float)]
    L.rn_terrain2d_batch.argtypes = [vp, C.POINTER(Params), C.POINTER(u32), i32, vp, sz, i32, vp]
    L.rn_terrain2d_batch_chunk.argtypes = [vp, C.POINTER(Params), i32]
    L.rn_lcg_draws_device.argtypes = [vp, u32, u64, u64, vp, i32, vp]
    L.rn_selftest_math.argtypes = [vp, C.POINTER(u64)]
    L.rn_terrain2d_cell_count.restype = u64
    L.rn_terrain2d_cell_count.argtypes = [i32]
    L.rn_terrain2d_cell_level_offset.restype = u64
    L.rn_terrain2d_cell_level_offset.argtypes = [i32]
    L.rn_terrain2d_cells.argtypes = [vp, C.POINTER(Params), vp, sz, i32, vp, i32, vp]
    L.rn_volume3d.argtypes = [vp, C.POINTER(VolumeParams), i32, i32, vp, vp]
    L.rn_volume3d_host.argtypes = [vp, C.POINTER(VolumeParams), i32, i32, vp]
    L.rn_volume3d_slab_elems.restype = u64
    L.rn_volume3d_slab_elems.argtypes = [i32, i32]
    L.rn_rock_draw_count.restype = C.c_uint64
    L.rn_rock_draw_count.argtypes = [i32]
    L.rn_rock.restype = i32
    L.rn_rock.argtypes = [vp, C.POINTER(RockParams), vp, vp, i32, vp, C.POINTER(C.c_uint32)]
    L.rn_rock_batch.restype = i32
    L.rn_rock_batch.argtypes = [vp, u32, i32, C.POINTER(C.c_int32), C.c_float, C.POINTER(vp), C.POINTER(vp), i32, vp, C.POINTER(C.c_int32),
                                C.POINTER(C.c_float), C.POINTER(u32), C.POINTER(u32)]
    L.rn_cloud_draw_count.restype = C.c_uint64
    L.rn_cloud_draw_count.argtypes = [i32]
    L.rn_cloud.restype = i32
    L.rn_cloud.argtypes = [vp, C.POINTER(RockParams), vp, vp, i32, vp, C.POINTER(C.c_uint32)]
    L.rn_device_alloc.argtypes = [vp, sz, C.POINTER(vp)]
    L.rn_device_free.argtypes = [vp, vp]
    L.rn_host_alloc.argtypes = [sz, C.POINTER(vp)]
    L.rn_host_free.argtypes = [vp]
    L.rn_copy_to_host.argtypes = [vp, vp, vp, sz]
    L.rn_create_multi.argtypes = [C.POINTER(i32), i32, C.POINTER(vp)]
    L.rn_destroy_multi.argtypes = [vp]
    L.rn_multi_size.argtypes = [vp]
    L.rn_multi_context.restype = vp
    L.rn_multi_context.argtypes = [vp, i32]
    L.rn_slab_rows.argtypes = [i32, i32]
    L.rn_slab_window.argtypes = [i32, i32, i32, C.POINTER(Window)]
    L.rn_terrain2d_sharded.argtypes = [vp, C.POINTER(Params), C.POINTER(vp), sz, vp, sz]
    L.rn_terrain2d_batch_sharded.argtypes = [vp, C.POINTER(Params), C.POINTER(u32), i32, C.POINTER(vp), sz]
    L.rn_gather.argtypes = [vp, C.POINTER(Params), C.POINTER(vp), C.POINTER(vp), i32]
    _lib = L
    return L


def default_params(**overrides) -> Params:
    p = Params()
    lib().rn_params_default(C.byref(p))
    for k, v in overrides.items():
        if not hasattr(p, k):
            raise AttributeError(k)
        setattr(p, k, v)
    return p


def _check(status: int):
    if status != RN_OK:
        raise RedNoiseError(status, (lib().rn_last_error_string() or b"").decode() or
                            lib().rn_status_string(status).decode())


def side(depth: int) -> int:
    return lib().rn_terrain2d_side(depth)


class Context:
    """rn_context: one per CUDA device (one process per GPU)."""

    def __init__(self, device: int = -1):
        self._h = C.c_void_p()
        _check(lib().rn_create(device, C.byref(self._h)))

    def close(self):
        if self._h:
            lib().rn_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    @property
    def launches(self) -> int:
        return int(lib().rn_launch_count(self._h))

    def set_algorithm(self, algo: int):
        _check(lib().rn_set_algorithm(self._h, algo))

    def set_pipelining(self, on: bool):
        _check(lib().rn_set_pipelining(self._h, 1 if on else 0))

    def set_profiling(self, on: bool):
        _check(lib().rn_set_profiling(self._h, 1 if on else 0))

    def last_stage_times(self):
        """Device ms of each launch of the last call (coarse to fine); needs set_profiling(True)."""
        buf = (C.c_float * 16)()
        n = lib().rn_last_stage_times(self._h, buf, 16)
        if n < 0:
            _check(-n)
        return [float(buf[k]) for k in range(n)]

    # -- raw pointer form (what the C ABI takes) ---------------------------------------
    def terrain2d_ptr(self, p: Params, ptr: int, *, window: Window | None = None, pitch_elems: int = 0,
                      dst_is_device: bool = True, stream: int = 0, want_volcano_pos: bool = False):
        vpos = (C.c_float * 3)() if want_volcano_pos else None
        _check(lib().rn_terrain2d(self._h, C.byref(p), C.byref(window) if window is not None else None,
                                  C.c_void_p(ptr), pitch_elems, 1 if dst_is_device else 0,
                                  C.c_void_p(stream) if stream else None, vpos))
        return tuple(vpos) if want_volcano_pos else None

    def batch_chunk(self, p: Params, ntiles: int) -> int:
        """Fields per chunk that rn_terrain2d_batch uses for a device destination (include/rednoise.h)."""
        return int(lib().rn_terrain2d_batch_chunk(self._h, C.byref(p), ntiles))

    def terrain2d_batch_ptr(self, p: Params, seeds, ptr: int, *, tile_stride_elems: int = 0,
                            dst_is_device: bool = True, stream: int = 0, seeds_ptr: int = 0):
        """seeds: sequence of ints, or -- with seeds_ptr -- the number of uint32 seeds at that HOST address (e.g. pinned memory,
        which must stay untouched until the stream has passed the call: the upload is then truly asynchronous)."""
        if seeds_ptr:
            n = int(seeds)
            arr = C.cast(C.c_void_p(seeds_ptr), C.POINTER(C.c_uint32))
        else:
            n = len(seeds)
            arr = (C.c_uint32 * n)(*[int(s) & 0xFFFFFFFF for s in seeds])
        _check(lib().rn_terrain2d_batch(self._h, C.byref(p), arr, n, C.c_void_p(ptr), tile_stride_elems,
                                        1 if dst_is_device else 0, C.c_void_p(stream) if stream else None))

    # -- numpy conveniences (host destination) -----------------------------------------
    def terrain2d_host(self, p: Params, window: Window | None = None, want_volcano_pos: bool = False):
        s = side(p.depth)
        ni, nj = (window.ni, window.nj) if window is not None else (s, s)
        out = np.empty((ni, nj, 4), np.float32)
        vp = self.terrain2d_ptr(p, out.ctypes.data, window=window, dst_is_device=False,
                                want_volcano_pos=want_volcano_pos)
        return (out, vp) if want_volcano_pos else out

    def terrain2d_batch_host(self, p: Params, seeds):
        s = side(p.depth)
        out = np.empty((len(seeds), s, s, 4), np.float32)
        self.terrain2d_batch_ptr(p, seeds, out.ctypes.data, dst_is_device=False)
        return out

    def cells_ptr(self, p: Params, cells_ptr: int, *, field_ptr: int = 0, field_pitch: int = 0, field_is_device: bool = True,
                  cells_is_device: bool = True, stream: int = 0):
        _check(lib().rn_terrain2d_cells(self._h, C.byref(p), C.c_void_p(field_ptr) if field_ptr else None, field_pitch,
                                        1 if field_is_device else 0, C.c_void_p(cells_ptr), 1 if cells_is_device else 0,
                                        C.c_void_p(stream) if stream else None))

    def cells_host(self, p: Params, field: np.ndarray | None = None) -> np.ndarray:
        """Cell-quadtree attributes (structured array, CELL_DTYPE) of `field` (host array) or of the field of p."""
        n = int(lib().rn_terrain2d_cell_count(p.depth))
        out = np.zeros(n, CELL_DTYPE)
        if field is not None:
            field = np.ascontiguousarray(field, np.float32)
        self.cells_ptr(p, out.ctypes.data, field_ptr=field.ctypes.data if field is not None else 0, field_is_device=False,
                       cells_is_device=False)
        return out

    def volume3d_ptr(self, seed: int, depth: int, amplitude: float, ptr: int, *, i0: int = 0, ni: int | None = None, stream: int = 0):
        """3D extension: planes [i0, i0+ni) of the (W+1)^3 float volume into the slab-sized device buffer at `ptr`
        (ni x (W+1) x (W+1) floats)."""
        vp_ = VolumeParams(seed & 0xFFFFFFFF, depth, amplitude)
        side = (1 << depth) + 1
        _check(lib().rn_volume3d(self._h, C.byref(vp_), i0, side - i0 if ni is None else ni, C.c_void_p(ptr),
                                 C.c_void_p(stream) if stream else None))

    def volume3d_to_host_ptr(self, seed: int, depth: int, amplitude: float, host_ptr: int, *, i0: int = 0, ni: int | None = None):
        """3D extension: the same planes into HOST memory at `host_ptr` (blocking; generation and D2H copy overlap chunk by chunk)."""
        vp_ = VolumeParams(seed & 0xFFFFFFFF, depth, amplitude)
        side = (1 << depth) + 1
        _check(lib().rn_volume3d_host(self._h, C.byref(vp_), i0, side - i0 if ni is None else ni, C.c_void_p(host_ptr)))

    def rock_ptr(self, state: int, depth: int, amplitude: float, faces_ptr: int, cells_ptr: int, *, dst_is_device: bool = True, stream: int = 0,
                 cloud: bool = False) -> int:
        """Rock generator (Rock.cpp; cloud=True: Cloud.cpp): faces / cells at the given pointers (0 = not wanted); returns the generator
        state afterwards."""
        rp = RockParams(state & 0xFFFFFFFF, depth, amplitude)
        after = C.c_uint32(0)
        fn = lib().rn_cloud if cloud else lib().rn_rock
        _check(fn(self._h, C.byref(rp), C.c_void_p(faces_ptr) if faces_ptr else None, C.c_void_p(cells_ptr) if cells_ptr else None,
                             1 if dst_is_device else 0, C.c_void_p(stream) if stream else None, C.byref(after)))
        return int(after.value)

    def rock_host(self, state: int, depth: int, amplitude: float = 0.3, want_cells: bool = True, cloud: bool = False):
        """(faces[6 or 2, W+1, W+1, 2] float32, cells[6 or 2, count] CELL_DTYPE or None, state_after) in host memory."""
        side = (1 << depth) + 1
        nf = 2 if cloud else 6
        faces = np.empty((nf, side, side, 2), np.float32)
        cells = np.zeros((nf, int(lib().rn_terrain2d_cell_count(depth))), CELL_DTYPE) if want_cells else None
        after = self.rock_ptr(state, depth, amplitude, faces.ctypes.data, cells.ctypes.data if want_cells else 0, dst_is_device=False, cloud=cloud)
        return faces, cells, after

    def rock_batch_host(self, state: int, nrocks: int, depths=None, amplitude: float = 0.3):
        """rn_rock_batch with host destinations: ([faces_i], [cells_i], depths, exponents, states_at_entry, state_after)."""
        nd = (C.c_int32 * nrocks)()
        ex = (C.c_float * nrocks)()
        st = (C.c_uint32 * nrocks)()
        after = C.c_uint32(0)
        if depths is None:
            # depths are decided by the stream: ask for them first (no destinations needed for the dry walk: use depth 12 sized? no --
            # the walk is host-only, so run it through the library's own helper by generating cells only on a second call)
            s = state & 0xFFFFFFFF
            dl = []
            for _ in range(nrocks):
                s = (s * 214013 + 2531011) & 0xFFFFFFFF
                r = (s >> 16) & 0x7FFF
                e = np.float32(-5.0) + np.float32(5.0) * (np.float32(r) / np.float32(32767.0))
                dl.append(int(np.float32(e + np.float32(7.0))))
                s = int(lib().rn_lcg_state_after(s, int(lib().rn_rock_draw_count(dl[-1]))))
        else:
            dl = [int(x) for x in depths]
        faces = [np.empty((6, (1 << d) + 1, (1 << d) + 1, 2), np.float32) for d in dl]
        cells = [np.zeros((6, int(lib().rn_terrain2d_cell_count(d))), CELL_DTYPE) for d in dl]
        fp = (C.c_void_p * nrocks)(*[f.ctypes.data for f in faces])
        cp = (C.c_void_p * nrocks)(*[c.ctypes.data for c in cells])
        darr = (C.c_int32 * nrocks)(*dl) if depths is not None else None
        _check(lib().rn_rock_batch(self._h, state & 0xFFFFFFFF, nrocks, darr, amplitude, fp, cp, 0, None, nd, ex, st, C.byref(after)))
        return faces, cells, list(nd), [float(x) for x in ex], [int(x) for x in st], int(after.value)

    def selftest_math(self):
        buf = (C.c_uint64 * 3)()
        _check(lib().rn_selftest_math(self._h, buf))
        return [int(x) for x in buf]

    def lcg_draws_host(self, seed: int, first: int, n: int) -> np.ndarray:
        out = np.empty(n, np.uint16)
        _check(lib().rn_lcg_draws_device(self._h, seed & 0xFFFFFFFF, first, n, C.c_void_p(out.ctypes.data), 0, None))
        return out


class Multi:
    """rn_multi: several devices driven from one host thread (include/rednoise.h, "several GPUs from one host thread")."""

    def __init__(self, devices):
        self._h = C.c_void_p()
        arr = (C.c_int * len(devices))(*devices)
        _check(lib().rn_create_multi(arr, len(devices), C.byref(self._h)))
        self.n = len(devices)

    def close(self):
        if self._h:
            lib().rn_destroy_multi(self._h)
            self._h = C.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def terrain2d_sharded(self, p: Params, slab_ptrs=None, slab_pitch: int = 0, host_ptr: int = 0, host_pitch: int = 0):
        arr = (C.c_void_p * self.n)(*slab_ptrs) if slab_ptrs is not None else None
        _check(lib().rn_terrain2d_sharded(self._h, C.byref(p), arr, slab_pitch, C.c_void_p(host_ptr) if host_ptr else None, host_pitch))

    def gather(self, p: Params, slab_ptrs, full_ptrs, use_nccl: bool = True):
        a = (C.c_void_p * self.n)(*slab_ptrs)
        b = (C.c_void_p * self.n)(*full_ptrs)
        _check(lib().rn_gather(self._h, C.byref(p), a, b, 1 if use_nccl else 0))


def slab_window(side: int, part: int, nparts: int) -> Window:
    w = Window()
    _check(lib().rn_slab_window(side, part, nparts, C.byref(w)))
    return w
