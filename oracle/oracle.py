"""ctypes front-end of the oracle -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / ``--impl reference`` legs may
import this module.  The product package must never do so.

Two libraries:
  * ``liboracle.so``       -- oracle/restate.cc, the CPU restatement (always buildable);
  * ``_ref/libref_*.so``   -- the REAL reference text compiled by oracle/Makefile (exists wherever
                              it was built while /root/reference was present; travels to the
                              GPU box as a prebuilt file).
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from ctypes import POINTER, byref, c_char_p, c_int, c_int32, c_int64, c_uint, c_uint64, c_void_p

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
INF = 0xFFFFFFFF

PATH_BIT = 0x2000
START_BIT = 0x4000
FINISH_BIT = 0x8000
BUILDER_BIT = 0x1000
PAINT_MASK = 0x00F0
CACHE_MASK = 0x0F00
MEASURE_BIT = 0x0200
HUNT, GATHER, CORNERS = 0, 1, 2


def build(verbose: bool = False) -> None:
    """Compile liboracle.so and, when /root/reference is present, oracle/_ref/."""
    out = None if verbose else subprocess.DEVNULL
    subprocess.check_call(["make", "-C", HERE, "all"], stdout=out)


def _p(a: np.ndarray):
    return a.ctypes.data_as(c_void_p)


def _load(path: str):
    return ctypes.CDLL(path, mode=getattr(os, "RTLD_LOCAL", 0))


_ORC = None
_REF_BFS = None
_REF_DIST = None


def orc():
    global _ORC
    if _ORC is None:
        path = os.path.join(HERE, "liboracle.so")
        if not os.path.exists(path):
            build()
        _ORC = _load(path)
    return _ORC


def read_maze_file(path: str) -> np.ndarray:
    """A raw maze file (include/labyrinth_b200.h: "LABMAZE1", int32 rows, cols, then uint16 LE row-major), read
    without the product's library so that oracle and product consume the same BYTES independently."""
    with open(path, "rb") as f:
        head = f.read(16)
        if len(head) != 16 or head[:8] != b"LABMAZE1":
            raise ValueError(f"{path} is not a maze file")
        rows, cols = (int(x) for x in np.frombuffer(head[8:], "<i4"))
        g = np.fromfile(f, dtype="<u2", count=rows * cols)
    if g.size != rows * cols:
        raise ValueError(f"{path} is truncated")
    return g.astype(np.uint16, copy=False).reshape(rows, cols)


def write_maze_file(path: str, grid: np.ndarray) -> None:
    g = np.ascontiguousarray(grid, "<u2")
    with open(path, "wb") as f:
        f.write(b"LABMAZE1" + np.asarray(g.shape, "<i4").tobytes())
        g.tofile(f)


def have_ref() -> bool:
    return os.path.exists(os.path.join(HERE, "_ref", "libref_bfs.so")) and os.path.exists(
        os.path.join(HERE, "_ref", "libref_dist.so")
    )


def ref_bfs():
    global _REF_BFS
    if _REF_BFS is None:
        _REF_BFS = _load(os.path.join(HERE, "_ref", "libref_bfs.so"))
    return _REF_BFS


_REF_RUNS = None


def ref_runs_lib():
    global _REF_RUNS
    if _REF_RUNS is None:
        _REF_RUNS = _load(os.path.join(HERE, "_ref", "libref_runs.so"))
    return _REF_RUNS


def have_ref_runs() -> bool:
    return os.path.exists(os.path.join(HERE, "_ref", "libref_runs.so"))


def ref_runs(grid: np.ndarray, seed: int = 0):
    """The real Runs::paint_runs (painters/runs.cc:130-161) -> (grid_after, runs u64 (2^64-1 = absent), max)"""
    rows, cols = grid.shape
    g = np.ascontiguousarray(grid, np.uint16).copy()
    r = np.zeros((rows, cols), np.uint64)
    mx = c_uint64()
    ref_runs_lib().ref_runs(c_int(rows), c_int(cols), _p(g), c_uint(seed), _p(r), byref(mx))
    return g, r, mx.value


def runs_map(grid: np.ndarray):
    """painters/runs.cc:130-161 restated -> (grid_after, runs u32 (INF = absent), max, settled)"""
    rows, cols = grid.shape
    g = np.ascontiguousarray(grid, np.uint16).copy()
    r = np.zeros((rows, cols), np.uint32)
    mx, st = c_uint64(), c_uint64()
    orc().orc_runs_map(c_int(rows), c_int(cols), _p(g), _p(r), byref(mx), byref(st))
    return g, r, mx.value, st.value


def ref_dist():
    global _REF_DIST
    if _REF_DIST is None:
        _REF_DIST = _load(os.path.join(HERE, "_ref", "libref_dist.so"))
    return _REF_DIST


# --------------------------------------------------------------------------- reference wrappers
MODS = {None: 0, "none": 0, "cross": 1, "x": 2}


def ref_build(builder: str, rows: int, cols: int, seed: int, mod=None, lib=None) -> np.ndarray:
    lib = lib or ref_bfs()
    g = np.zeros((rows, cols), np.uint16)
    rc = lib.ref_build(c_char_p(builder.encode()), c_int(rows), c_int(cols), c_uint(seed), c_int(MODS[mod]), _p(g))
    if rc != 0:
        raise ValueError(f"unknown builder {builder}")
    return g


def ref_distance(grid: np.ndarray, seed: int = 0):
    """-> (grid_after, dist u64 (2^64-1 = absent), max, rgb int32 [R,C,3])"""
    rows, cols = grid.shape
    g = np.ascontiguousarray(grid, np.uint16).copy()
    d = np.zeros((rows, cols), np.uint64)
    rgb = np.zeros((rows, cols, 3), np.int32)
    mx = c_uint64()
    ref_dist().ref_distance(c_int(rows), c_int(cols), _p(g), c_uint(seed), _p(d), byref(mx), _p(rgb))
    return g, d, mx.value, rgb


def _solver_sequential(lib, fname, game, grid, starts, order, max_path=None):
    rows, cols = grid.shape
    g = np.ascontiguousarray(grid, np.uint16).copy()
    max_path = max_path or rows * cols
    st = np.ascontiguousarray(starts, np.int32).reshape(4, 2)
    od = np.ascontiguousarray(order, np.int32)
    paths = np.zeros((4, max_path, 2), np.int32)
    lens = np.zeros(4, np.int64)
    win = c_int32()
    getattr(lib, fname)(
        c_int(game), c_int(rows), c_int(cols), _p(g), _p(st), _p(od), c_int(len(od)), _p(paths),
        c_int64(max_path), _p(lens), byref(win),
    )
    return g, [paths[i, : lens[i]].copy() for i in range(4)], win.value


def ref_solver_sequential(game, grid, starts, order=(0, 1, 2, 3)):
    return _solver_sequential(ref_bfs(), "ref_solver_sequential", game, grid, starts, order)


def ref_game(game: int, grid: np.ndarray, seed: int) -> np.ndarray:
    rows, cols = grid.shape
    g = np.ascontiguousarray(grid, np.uint16).copy()
    ref_bfs().ref_game(c_int(game), c_int(rows), c_int(cols), _p(g), c_uint(seed))
    return g


def ref_pick_random_point(grid: np.ndarray, seed: int):
    rows, cols = grid.shape
    g = np.ascontiguousarray(grid, np.uint16)
    out = np.zeros(2, np.int32)
    ref_bfs().ref_pick_random_point(c_int(rows), c_int(cols), _p(g), c_uint(seed), _p(out))
    return int(out[0]), int(out[1])


def ref_set_corner_starts(grid: np.ndarray):
    rows, cols = grid.shape
    g = np.ascontiguousarray(grid, np.uint16)
    out = np.zeros(8, np.int32)
    ref_bfs().ref_set_corner_starts(c_int(rows), c_int(cols), _p(g), _p(out))
    return out.reshape(4, 2)


# ------------------------------------------------------------------------- restatement wrappers
def distance_map(grid: np.ndarray):
    """painters/distance.cc:129-153 restated -> (grid_after, dist u32, max, settled)"""
    rows, cols = grid.shape
    g = np.ascontiguousarray(grid, np.uint16).copy()
    d = np.zeros((rows, cols), np.uint32)
    mx, st = c_uint64(), c_uint64()
    orc().orc_distance_map(c_int(rows), c_int(cols), _p(g), _p(d), byref(mx), byref(st))
    return g, d, mx.value, st.value


def painter_rgb(dist: np.ndarray, mx: int, color_choice: int) -> np.ndarray:
    rows, cols = dist.shape
    rgb = np.zeros((rows, cols, 3), np.int32)
    orc().orc_painter_rgb(c_int(rows), c_int(cols), _p(np.ascontiguousarray(dist, np.uint32)), c_uint64(mx),
                          c_int(color_choice), _p(rgb))
    return rgb


def bfs_levels(grid: np.ndarray, sr: int, sc: int) -> np.ndarray:
    rows, cols = grid.shape
    d = np.zeros((rows, cols), np.uint32)
    orc().orc_bfs_levels(c_int(rows), c_int(cols), _p(np.ascontiguousarray(grid, np.uint16)), c_int(sr), c_int(sc), _p(d))
    return d


def solver_sequential(game, grid, starts, order=(0, 1, 2, 3)):
    return _solver_sequential(orc(), "orc_solver_sequential", game, grid, starts, order)


def pick_random_point(grid: np.ndarray, seed: int):
    rows, cols = grid.shape
    out = np.zeros(2, np.int32)
    orc().orc_pick_random_point(c_int(rows), c_int(cols), _p(np.ascontiguousarray(grid, np.uint16)), c_uint(seed), _p(out))
    return int(out[0]), int(out[1])


def set_corner_starts(grid: np.ndarray):
    rows, cols = grid.shape
    out = np.zeros(8, np.int32)
    orc().orc_set_corner_starts(c_int(rows), c_int(cols), _p(np.ascontiguousarray(grid, np.uint16)), _p(out))
    return out.reshape(4, 2)


def place_hunt(grid: np.ndarray, seed: int):
    rows, cols = grid.shape
    g = np.ascontiguousarray(grid, np.uint16).copy()
    s, f = np.zeros(2, np.int32), np.zeros(2, np.int32)
    rc = orc().orc_place_hunt(c_int(rows), c_int(cols), _p(g), c_uint(seed), _p(s), _p(f))
    assert rc == 0
    return g, s, f


def place_gather(grid: np.ndarray, seed: int):
    rows, cols = grid.shape
    g = np.ascontiguousarray(grid, np.uint16).copy()
    s, f = np.zeros(2, np.int32), np.zeros(8, np.int32)
    rc = orc().orc_place_gather(c_int(rows), c_int(cols), _p(g), c_uint(seed), _p(s), _p(f))
    assert rc == 0
    return g, s, f.reshape(4, 2)


def place_corners(grid: np.ndarray, seed: int):
    rows, cols = grid.shape
    g = np.ascontiguousarray(grid, np.uint16).copy()
    s, f = np.zeros(8, np.int32), np.zeros(2, np.int32)
    rc = orc().orc_place_corners(c_int(rows), c_int(cols), _p(g), c_uint(seed), _p(s), _p(f))
    assert rc == 0
    return g, s.reshape(4, 2), f


def canon_hunt(grid: np.ndarray, start, finish):
    """-> (grid_after, winner, path[n,2])"""
    rows, cols = grid.shape
    g = np.ascontiguousarray(grid, np.uint16).copy()
    mp = rows * cols
    path = np.zeros((mp, 2), np.int32)
    n, w = c_int64(), c_int32()
    orc().orc_canon_hunt(c_int(rows), c_int(cols), _p(g), _p(np.ascontiguousarray(start, np.int32)),
                         _p(np.ascontiguousarray(finish, np.int32)), byref(w), _p(path), c_int64(mp), byref(n))
    return g, w.value, path[: n.value].copy()


def canon_gather(grid: np.ndarray, start, finishes):
    """-> (grid_after, claimed_by[4], paths[4])"""
    rows, cols = grid.shape
    g = np.ascontiguousarray(grid, np.uint16).copy()
    mp = rows * cols
    paths = np.zeros((4, mp, 2), np.int32)
    lens = np.zeros(4, np.int64)
    claimed = np.zeros(4, np.int32)
    orc().orc_canon_gather(c_int(rows), c_int(cols), _p(g), _p(np.ascontiguousarray(start, np.int32)),
                           _p(np.ascontiguousarray(finishes, np.int32)), _p(claimed), _p(paths), c_int64(mp), _p(lens))
    return g, claimed, [paths[i, : lens[i]].copy() for i in range(4)]


def canon_corners(grid: np.ndarray, starts, finish):
    """-> (grid_after, winner, path[n,2])"""
    rows, cols = grid.shape
    g = np.ascontiguousarray(grid, np.uint16).copy()
    mp = rows * cols
    path = np.zeros((mp, 2), np.int32)
    n, w = c_int64(), c_int32()
    orc().orc_canon_corners(c_int(rows), c_int(cols), _p(g), _p(np.ascontiguousarray(starts, np.int32)),
                            _p(np.ascontiguousarray(finish, np.int32)), byref(w), _p(path), c_int64(mp), byref(n))
    return g, w.value, path[: n.value].copy()
