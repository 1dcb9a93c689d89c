"""labyrinth_b200 -- Python plumbing over liblabyrinth_b200.so (the B200 BFS hot path).

The product is the CUDA library behind ``include/labyrinth_b200.h``; this module only loads it
with ctypes and gives tests/bench a convenient handle.  PyTorch is used by callers for device
memory, streams and ``torch.distributed`` -- nothing in here falls back to the CPU: if the shared
library is missing, or no CUDA device is present when a search is asked for, it raises.

The directory name contains a hyphen (it is the name the task prescribes), so import it with
``__graft_entry__.load_package()`` or ``tests/conftest.py``'s ``lab`` fixture.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, Structure, byref, c_char_p, c_float, c_int, c_int32, c_size_t, c_uint16, c_uint32, c_uint64, c_void_p
from dataclasses import dataclass
from typing import Optional, Sequence

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("LAB_LIB") or os.path.join(HERE, "liblabyrinth_b200.so")
INF = 0xFFFFFFFF

# Square bits: maze/maze.cc:170-173, solvers/solve_utilities.cc:59-79, painters/rgb.cc:21-22
PATH_BIT, START_BIT, FINISH_BIT, BUILDER_BIT = 0x2000, 0x4000, 0x8000, 0x1000
PAINT_MASK, CACHE_MASK, MEASURE_BIT = 0x00F0, 0x0F00, 0x0200

#: every symbol include/labyrinth_b200.h declares
ABI_SYMBOLS = (
    "lab_grid_create", "lab_grid_create_on_device", "lab_grid_upload", "lab_grid_download", "lab_grid_destroy",
    "lab_grid_set_stream", "lab_grid_device_squares", "lab_grid_rows", "lab_grid_cols", "lab_grid_or_bits",
    "lab_grid_and_bits", "lab_distance_map", "lab_distance_start", "lab_levels_device", "lab_levels_bind", "lab_levels_download",
    "lab_distance_paint", "lab_pixels_device", "lab_pixels_bind", "lab_runs_map", "lab_runs_paint",
    "lab_bfs_hunt", "lab_bfs_gather", "lab_bfs_corners", "lab_last_painted", "lab_last_timing", "lab_last_stats",
    "lab_last_error", "lab_pick_random_point", "lab_set_corner_starts", "lab_place_hunt", "lab_place_gather",
    "lab_place_corners", "lab_build_maze", "lab_device_info", "lab_maze_write", "lab_maze_header", "lab_maze_read",
    "lab_band_cols_padded", "lab_band_begin", "lab_band_path_rows", "lab_band_set_neighbour_paths", "lab_band_relax",
    "lab_band_export_edges", "lab_band_import_edges", "lab_band_finish",
    "lab_band_tree_bind", "lab_band_tree_chunk_words", "lab_band_tree_count", "lab_band_tree_rank_local", "lab_band_tree_flood",
    "lab_band_tree_prefix_local", "lab_band_tree_levels", "lab_band_room",
    "lab_distance_map_sharded", "lab_sharded_last_error", "lab_band_begin_lattice", "lab_band_lat_super_rows", "lab_band_lat_words", "lab_band_lat_bind", "lab_band_lat_set_peers", "lab_band_lat_link", "lab_band_lat_flood",
    "lab_band_lat_levels",
    "lab_band_begin_game", "lab_band_level_at", "lab_band_paint", "lab_band_walk", "lab_band_recolour", "lab_band_finish_game",
    "lab_band_begin_corners", "lab_band_planes", "lab_band_level_at_plane", "lab_band_paint_planes",
    "lab_band_lat_path_rank", "lab_band_lat_path_mark", "lab_grid_set_option",
)


OPT_SOLVER_LEVELS = 1  # lab_grid_set_option


class LabError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"liblabyrinth_b200 error {code}: {msg}")
        self.code = code


class Point(Structure):
    _fields_ = [("row", c_int32), ("col", c_int32)]


class Timing(Structure):
    _fields_ = [(n, c_float) for n in ("h2d_ms", "setup_ms", "pack_ms", "search_ms", "post_ms", "total_ms", "d2h_ms")]

    def as_dict(self):
        return {n: float(getattr(self, n)) for n, _ in self._fields_}


class Stats(Structure):
    _fields_ = [(n, c_uint64) for n in ("tiles", "activations", "empty", "levels", "settled", "rechecks", "resettles",
                                        "queue_pushes", "direct")] + [("worker_warps", c_uint32), ("error", c_uint32)] + [
        (n, c_uint64) for n in ("cyc_load", "cyc_loop", "cyc_publish", "cyc_idle", "cyc_step", "cyc_emit", "batches",
                                "tree_squares", "tree_pairs", "tree_crossings", "tree_used", "launches")]

    def as_dict(self):
        return {n: int(getattr(self, n)) for n, _ in self._fields_}


_lib = None


def lib():
    """The loaded shared library; raises if it has not been built (no fallback)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(nvcc, sm_100a).  There is no CPU fallback."
            )
        L = ctypes.CDLL(LIB_PATH)
        L.lab_last_error.restype = c_char_p
        L.lab_grid_device_squares.restype = c_void_p
        L.lab_levels_device.restype = c_void_p
        L.lab_pixels_device.restype = c_void_p
        L.lab_last_painted.restype = c_uint64
        L.lab_distance_start.restype = Point
        L.lab_grid_create.argtypes = [c_int, c_int, c_void_p, POINTER(c_void_p)]
        L.lab_grid_create_on_device.argtypes = [c_int, c_int, c_void_p, POINTER(c_void_p)]
        L.lab_grid_upload.argtypes = [c_void_p, c_void_p]
        L.lab_grid_download.argtypes = [c_void_p, c_void_p]
        L.lab_grid_destroy.argtypes = [c_void_p]
        L.lab_grid_set_stream.argtypes = [c_void_p, c_void_p]
        L.lab_grid_device_squares.argtypes = [c_void_p]
        L.lab_grid_or_bits.argtypes = [c_void_p, c_void_p, c_void_p, c_int]
        L.lab_grid_and_bits.argtypes = [c_void_p, c_uint16]
        L.lab_distance_map.argtypes = [c_void_p, c_int, c_int, c_void_p, POINTER(c_uint64), POINTER(c_uint64)]
        L.lab_distance_start.argtypes = [c_int, c_int]
        L.lab_levels_device.argtypes = [c_void_p, c_int]
        L.lab_levels_bind.argtypes = [c_void_p, c_int, c_void_p]
        L.lab_distance_paint.argtypes = [c_void_p, c_int, c_void_p]
        L.lab_runs_map.argtypes = [c_void_p, c_int, c_int, c_void_p, POINTER(c_uint64), POINTER(c_uint64)]
        L.lab_runs_paint.argtypes = [c_void_p, c_int, c_void_p]
        L.lab_pixels_device.argtypes = [c_void_p]
        L.lab_pixels_bind.argtypes = [c_void_p, c_void_p]
        L.lab_bfs_hunt.argtypes = [c_void_p, Point, Point, POINTER(c_int32), c_void_p, c_uint64, POINTER(c_uint64)]
        L.lab_bfs_gather.argtypes = [c_void_p, Point, c_void_p, c_void_p, c_void_p, c_uint64, c_void_p]
        L.lab_bfs_corners.argtypes = [c_void_p, c_void_p, Point, POINTER(c_int32), c_void_p, c_uint64, POINTER(c_uint64)]
        L.lab_last_painted.argtypes = [c_void_p]
        L.lab_last_timing.argtypes = [c_void_p, POINTER(Timing)]
        L.lab_last_stats.argtypes = [c_void_p, POINTER(Stats)]
        L.lab_pick_random_point.argtypes = [c_int, c_int, c_void_p, c_uint32, POINTER(Point)]
        L.lab_set_corner_starts.argtypes = [c_int, c_int, c_void_p, c_void_p]
        L.lab_place_hunt.argtypes = [c_int, c_int, c_void_p, c_uint32, POINTER(Point), POINTER(Point)]
        L.lab_place_gather.argtypes = [c_int, c_int, c_void_p, c_uint32, POINTER(Point), c_void_p]
        L.lab_place_corners.argtypes = [c_int, c_int, c_void_p, c_uint32, c_void_p, POINTER(Point)]
        L.lab_build_maze.argtypes = [c_char_p, c_char_p, c_int, c_int, c_uint32, c_void_p]
        L.lab_device_info.argtypes = [c_char_p, c_size_t]
        L.lab_maze_write.argtypes = [c_char_p, c_int, c_int, c_void_p]
        L.lab_maze_header.argtypes = [c_char_p, POINTER(c_int), POINTER(c_int)]
        L.lab_maze_read.argtypes = [c_char_p, c_int, c_int, c_void_p]
        L.lab_band_cols_padded.argtypes = [c_void_p]
        L.lab_band_begin.argtypes = [c_void_p, c_int, c_int]
        L.lab_band_path_rows.argtypes = [c_void_p, c_void_p, c_void_p]
        L.lab_band_set_neighbour_paths.argtypes = [c_void_p, c_void_p, c_void_p]
        L.lab_band_relax.argtypes = [c_void_p]
        L.lab_band_export_edges.argtypes = [c_void_p, c_void_p, c_void_p]
        L.lab_band_import_edges.argtypes = [c_void_p, c_void_p, c_void_p, POINTER(c_uint32)]
        L.lab_band_finish.argtypes = [c_void_p, POINTER(c_uint64), POINTER(c_uint64)]
        _lib = L
    return _lib


def _check(rc: int) -> None:
    if rc != 0:
        raise LabError(rc, lib().lab_last_error().decode(errors="replace"))


def _np_ptr(a: np.ndarray):
    return a.ctypes.data_as(c_void_p)


def device_info() -> tuple[int, str]:
    buf = ctypes.create_string_buffer(256)
    n = lib().lab_device_info(buf, 256)
    return n, buf.value.decode()


# ------------------------------------------------------------------------------ host-side helpers
def build_maze(builder: str, rows: int, cols: int, seed: int, mod: Optional[str] = None) -> np.ndarray:
    """Seeded linear-time builder (arena|rdfs|grid|kruskal|wilson, mod none|cross|x) -> uint16[rows, cols]."""
    g = np.empty((rows, cols), np.uint16)
    _check(lib().lab_build_maze(builder.encode(), (mod or "none").encode(), rows, cols, seed & 0xFFFFFFFF, _np_ptr(g)))
    return g


# The raw maze file (include/labyrinth_b200.h): "LABMAZE1", int32 rows, int32 cols, rows*cols uint16 LE, row-major.
MAZE_FILE_HEADER_BYTES = 16


def write_maze(path: str, squares: np.ndarray) -> None:
    g = np.ascontiguousarray(squares, np.uint16)
    _check(lib().lab_maze_write(os.fsencode(path), g.shape[0], g.shape[1], _np_ptr(g)))


def maze_header(path: str) -> tuple[int, int]:
    r, c = c_int(), c_int()
    _check(lib().lab_maze_header(os.fsencode(path), byref(r), byref(c)))
    return r.value, c.value


def read_maze(path: str, mmap: bool = False) -> np.ndarray:
    """-> uint16[rows, cols]; mmap=True maps the file read-only instead of copying it (the data starts 16-byte aligned)."""
    rows, cols = maze_header(path)
    if mmap:
        return np.memmap(path, dtype=np.uint16, mode="r", offset=MAZE_FILE_HEADER_BYTES, shape=(rows, cols))
    g = np.empty((rows, cols), np.uint16)
    _check(lib().lab_maze_read(os.fsencode(path), rows, cols, _np_ptr(g)))
    return g


def cached_maze(builder: str, rows: int, cols: int, seed: int, mod: Optional[str] = None, cache_dir: Optional[str] = None,
                mmap: bool = False) -> np.ndarray:
    """build_maze() through a file cache keyed by (builder, mod, size, seed): a 32767 x 32767 Kruskal maze takes two
    minutes of host time to build and a second to read back.  The cache lives in $LAB_MAZE_CACHE (default
    /dev/shm/lab_mazes, else the temp dir); several ranks may race for one entry (files are written under a
    temporary name and renamed)."""
    d = cache_dir or os.environ.get("LAB_MAZE_CACHE") or ("/dev/shm/lab_mazes" if os.path.isdir("/dev/shm") else os.path.join(
        __import__("tempfile").gettempdir(), "lab_mazes"))
    path = os.path.join(d, f"{builder}-{mod or 'none'}-{rows}x{cols}-{seed & 0xFFFFFFFF:08x}.labmaze")
    if os.path.exists(path):
        try:
            if maze_header(path) == (rows, cols):
                return read_maze(path, mmap=mmap)
        except LabError:
            pass
    g = build_maze(builder, rows, cols, seed, mod)
    try:
        os.makedirs(d, exist_ok=True)
        write_maze(path, g)
    except (OSError, LabError):
        pass  # a cache that cannot be written is only slower
    return g


def distance_map_sharded(squares: np.ndarray, devices: Sequence[int], start: Optional[tuple] = None, want_dist: bool = True):
    """lab_distance_map_sharded: one process, `devices` GPUs, row bands (peer stores over NVLink, no NCCL).
    -> (squares after, DistanceResult, devices used)"""
    g = np.ascontiguousarray(squares, np.uint16).copy()
    rows, cols = g.shape
    sr, sc = start if start is not None else distance_start(rows, cols)
    dist = np.empty((rows, cols), np.uint32) if want_dist else None
    mx, st, used = c_uint64(), c_uint64(), c_int()
    dev = (c_int * len(devices))(*devices)
    L = lib()
    L.lab_sharded_last_error.restype = c_char_p
    rc = L.lab_distance_map_sharded(len(devices), dev, rows, cols, _np_ptr(g), sr, sc, _np_ptr(dist) if want_dist else None, byref(mx), byref(st), byref(used))
    if rc != 0:
        raise LabError(rc, (L.lab_sharded_last_error() or L.lab_last_error()).decode(errors="replace"))
    return g, DistanceResult(mx.value, st.value, dist), used.value


def distance_start(rows: int, cols: int) -> tuple[int, int]:
    p = lib().lab_distance_start(rows, cols)
    return p.row, p.col


def pick_random_point(grid: np.ndarray, seed: int) -> tuple[int, int]:
    p = Point()
    g = np.ascontiguousarray(grid, np.uint16)
    _check(lib().lab_pick_random_point(g.shape[0], g.shape[1], _np_ptr(g), seed & 0xFFFFFFFF, byref(p)))
    return p.row, p.col


def set_corner_starts(grid: np.ndarray) -> np.ndarray:
    g = np.ascontiguousarray(grid, np.uint16)
    out = np.zeros((4, 2), np.int32)
    _check(lib().lab_set_corner_starts(g.shape[0], g.shape[1], _np_ptr(g), _np_ptr(out)))
    return out


def place_hunt(grid: np.ndarray, seed: int):
    """-> (grid with start/finish bits, start (r,c), finish (r,c))   [Bfs::hunt :245-251]"""
    g = np.ascontiguousarray(grid, np.uint16).copy()
    s, f = Point(), Point()
    _check(lib().lab_place_hunt(g.shape[0], g.shape[1], _np_ptr(g), seed & 0xFFFFFFFF, byref(s), byref(f)))
    return g, (s.row, s.col), (f.row, f.col)


def place_gather(grid: np.ndarray, seed: int):
    g = np.ascontiguousarray(grid, np.uint16).copy()
    s = Point()
    f = np.zeros((4, 2), np.int32)
    _check(lib().lab_place_gather(g.shape[0], g.shape[1], _np_ptr(g), seed & 0xFFFFFFFF, byref(s), _np_ptr(f)))
    return g, (s.row, s.col), f


def place_corners(grid: np.ndarray, seed: int):
    g = np.ascontiguousarray(grid, np.uint16).copy()
    s = np.zeros((4, 2), np.int32)
    f = Point()
    _check(lib().lab_place_corners(g.shape[0], g.shape[1], _np_ptr(g), seed & 0xFFFFFFFF, _np_ptr(s), byref(f)))
    return g, s, (f.row, f.col)


# ------------------------------------------------------------------------------------ device grid
@dataclass
class DistanceResult:
    max: int
    settled: int
    dist: Optional[np.ndarray]


@dataclass
class SolveResult:
    winner: int
    claimed_by: Optional[np.ndarray]
    path_len: list
    paths: Optional[list]
    painted: int


class Grid:
    """A maze resident in one GPU's HBM (lab_grid).  Not thread-safe; one search at a time."""

    def __init__(self, handle: c_void_p, rows: int, cols: int, keepalive=None):
        self._h = handle
        self.rows, self.cols = rows, cols
        self._keep = keepalive

    @classmethod
    def from_host(cls, squares: np.ndarray) -> "Grid":
        g = np.ascontiguousarray(squares, np.uint16)
        h = c_void_p()
        _check(lib().lab_grid_create(g.shape[0], g.shape[1], _np_ptr(g), byref(h)))
        return cls(h, g.shape[0], g.shape[1])

    @classmethod
    def empty(cls, rows: int, cols: int) -> "Grid":
        h = c_void_p()
        _check(lib().lab_grid_create(rows, cols, None, byref(h)))
        return cls(h, rows, cols)

    @classmethod
    def from_device_ptr(cls, rows: int, cols: int, ptr: int, keepalive=None) -> "Grid":
        """Borrow caller-owned device memory (e.g. ``tensor.data_ptr()`` of an int16 CUDA tensor)."""
        h = c_void_p()
        _check(lib().lab_grid_create_on_device(rows, cols, c_void_p(ptr), byref(h)))
        return cls(h, rows, cols, keepalive)

    def close(self) -> None:
        if self._h:
            lib().lab_grid_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def set_stream(self, cuda_stream: int) -> None:
        _check(lib().lab_grid_set_stream(self._h, c_void_p(cuda_stream)))

    def upload(self, squares: np.ndarray) -> None:
        g = np.ascontiguousarray(squares, np.uint16)
        assert g.shape == (self.rows, self.cols)
        _check(lib().lab_grid_upload(self._h, _np_ptr(g)))

    def upload_ptr(self, host_ptr: int) -> None:
        _check(lib().lab_grid_upload(self._h, c_void_p(host_ptr)))

    def download(self, out: Optional[np.ndarray] = None) -> np.ndarray:
        if out is None:
            out = np.empty((self.rows, self.cols), np.uint16)
        _check(lib().lab_grid_download(self._h, _np_ptr(out)))
        return out

    def download_ptr(self, host_ptr: int) -> None:
        _check(lib().lab_grid_download(self._h, c_void_p(host_ptr)))

    def device_squares(self) -> int:
        return int(lib().lab_grid_device_squares(self._h) or 0)

    def levels(self, plane: int = 0) -> np.ndarray:
        """The level field of the last search (lab_levels_download); LabError if that search left none."""
        out = np.empty((self.rows, self.cols), np.uint32)
        _check(lib().lab_levels_download(self._h, c_int(plane), _np_ptr(out)))
        return out

    def levels_device(self, plane: int = 0) -> int:
        return int(lib().lab_levels_device(self._h, plane) or 0)

    def bind_levels(self, plane: int, ptr: int) -> None:
        _check(lib().lab_levels_bind(self._h, plane, c_void_p(ptr)))

    def or_bits(self, cells: Sequence[Sequence[int]], bits: Sequence[int]) -> None:
        c = np.ascontiguousarray(cells, np.int32).reshape(-1, 2)
        b = np.ascontiguousarray(bits, np.uint16)
        _check(lib().lab_grid_or_bits(self._h, _np_ptr(c), _np_ptr(b), len(b)))

    def and_bits(self, mask: int) -> None:
        _check(lib().lab_grid_and_bits(self._h, mask))

    # -- Distance::paint_distance_from_center, BFS part
    def distance_map(self, start: Optional[tuple] = None, want_dist: bool = True, dist_host_ptr: Optional[int] = None) -> DistanceResult:
        sr, sc = start if start is not None else distance_start(self.rows, self.cols)
        mx, st = c_uint64(), c_uint64()
        dist = None
        ptr = None
        if dist_host_ptr is not None:
            ptr = c_void_p(dist_host_ptr)
        elif want_dist:
            dist = np.empty((self.rows, self.cols), np.uint32)
            ptr = _np_ptr(dist)
        _check(lib().lab_distance_map(self._h, sr, sc, ptr, byref(mx), byref(st)))
        return DistanceResult(mx.value, st.value, dist)

    # -- Distance painter (painters/distance.cc:45-70): one R | G<<8 | B<<16 | 0xFF<<24 pixel per map square, 0 = wall glyph
    def paint(self, channel: int, want_pixels: bool = True, pixels_host_ptr: Optional[int] = None) -> Optional[np.ndarray]:
        px = None
        ptr = None
        if pixels_host_ptr is not None:
            ptr = c_void_p(pixels_host_ptr)
        elif want_pixels:
            px = np.empty((self.rows, self.cols), np.uint32)
            ptr = _np_ptr(px)
        _check(lib().lab_distance_paint(self._h, channel, ptr))
        return px

    # -- Runs::paint_runs (painters/runs.cc:130-161): straight-run lengths along the search's tree paths (one tree only)
    def runs_map(self, start: Optional[tuple] = None, want_runs: bool = True) -> DistanceResult:
        sr, sc = start if start is not None else distance_start(self.rows, self.cols)
        mx, st = c_uint64(), c_uint64()
        runs = np.empty((self.rows, self.cols), np.uint32) if want_runs else None
        _check(lib().lab_runs_map(self._h, sr, sc, _np_ptr(runs) if want_runs else None, byref(mx), byref(st)))
        return DistanceResult(mx.value, st.value, runs)

    def paint_runs(self, channel: int) -> np.ndarray:
        px = np.empty((self.rows, self.cols), np.uint32)
        _check(lib().lab_runs_paint(self._h, channel, _np_ptr(px)))
        return px

    def pixels_device(self) -> int:
        return int(lib().lab_pixels_device(self._h) or 0)

    def bind_pixels(self, ptr: int) -> None:
        _check(lib().lab_pixels_bind(self._h, c_void_p(ptr)))

    # -- Bfs::hunt / gather / corners, thread part
    def set_option(self, option: int, value: int):
        """lab_grid_set_option (OPT_SOLVER_LEVELS = 1: keep the solvers' level fields, the default; 0: gather without paths may skip its field)"""
        _check(lib().lab_grid_set_option(self._h, c_int(option), c_int(value)))

    def hunt(self, start, finish, want_path: bool = True) -> SolveResult:
        cap = self.rows * self.cols if want_path else 0
        path = np.empty((cap, 2), np.int32) if want_path else None
        w, n = c_int32(), c_uint64()
        _check(lib().lab_bfs_hunt(self._h, Point(*start), Point(*finish), byref(w), _np_ptr(path) if want_path else None, cap, byref(n)))
        return SolveResult(w.value, None, [n.value], [path[: n.value].copy()] if want_path else None, int(lib().lab_last_painted(self._h)))

    def gather(self, start, finishes, want_paths: bool = True) -> SolveResult:
        f = np.ascontiguousarray(finishes, np.int32).reshape(4, 2)
        cap = self.rows * self.cols if want_paths else 0
        claimed = np.zeros(4, np.int32)
        lens = np.zeros(4, np.uint64)
        bufs, ptrs = None, None
        if want_paths:
            bufs = [np.empty((cap, 2), np.int32) for _ in range(4)]
            ptrs = (c_void_p * 4)(*[b.ctypes.data for b in bufs])
        _check(lib().lab_bfs_gather(self._h, Point(*start), _np_ptr(f), _np_ptr(claimed), ptrs, cap, _np_ptr(lens)))
        paths = [bufs[k][: int(lens[k])].copy() for k in range(4)] if want_paths else None
        return SolveResult(-1, claimed, [int(x) for x in lens], paths, int(lib().lab_last_painted(self._h)))

    def corners(self, starts, finish, want_path: bool = True) -> SolveResult:
        s = np.ascontiguousarray(starts, np.int32).reshape(4, 2)
        cap = self.rows * self.cols if want_path else 0
        path = np.empty((cap, 2), np.int32) if want_path else None
        w, n = c_int32(), c_uint64()
        _check(lib().lab_bfs_corners(self._h, _np_ptr(s), Point(*finish), byref(w), _np_ptr(path) if want_path else None, cap, byref(n)))
        return SolveResult(w.value, None, [n.value], [path[: n.value].copy()] if want_path else None, int(lib().lab_last_painted(self._h)))

    def timing(self) -> dict:
        t = Timing()
        _check(lib().lab_last_timing(self._h, byref(t)))
        return t.as_dict()

    def stats(self) -> dict:
        s = Stats()
        _check(lib().lab_last_stats(self._h, byref(s)))
        return s.as_dict()
