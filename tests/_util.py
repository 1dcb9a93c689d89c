"""Shared helpers for the test suite (test infrastructure)."""
import ctypes
import glob
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = sorted(glob.glob(os.path.join(ROOT, "tests", "golden", "*.npz")))
MODS = {0: None, 1: "cross", 2: "x"}


def golden_id(path):
    return os.path.basename(path)[:-4]


def load_golden(path):
    z = np.load(path)
    rows, cols, seed, mod = (int(x) for x in z["meta"])
    builder = os.path.basename(path).split("_")[0]
    return z, builder, rows, cols, seed, MODS[mod]


def paths_equal(a, b):
    return len(a) == len(b) and all(np.asarray(x).shape == np.asarray(y).shape and (np.asarray(x) == np.asarray(y)).all() for x, y in zip(a, b))


def valid_shortest_path(grid, path, start, finish, length):
    """A reference-style path: finish side first, start last, finish excluded, wall respecting."""
    path = np.asarray(path).reshape(-1, 2)
    if len(path) != length:
        return False
    if length == 0:
        return True
    if tuple(path[-1]) != tuple(start):
        return False
    prev = np.asarray(finish)
    for p in path:
        if abs(int(p[0]) - int(prev[0])) + abs(int(p[1]) - int(prev[1])) != 1:
            return False
        if not (grid[p[0], p[1]] & 0x2000) and tuple(p) != tuple(start):
            return False
        prev = p
    return True


class Emu:
    """The product's kernel bodies run on the CPU warp emulator (tests/simt/libsimt_emu.so)."""

    def __init__(self):
        path = os.path.join(ROOT, "tests", "simt", "libsimt_emu.so")
        if not os.path.exists(path):
            import subprocess

            subprocess.check_call(["make", "-C", os.path.dirname(path)], stdout=subprocess.DEVNULL)
        self.lib = ctypes.CDLL(path)

    @staticmethod
    def _p(a):
        return a.ctypes.data_as(ctypes.c_void_p)

    def distance(self, grid, start=None, nwarps=4, sched=0, misalign=0):
        rows, cols = grid.shape
        # misalign: the Square array starts that many squares past a 16-byte boundary (a caller's sub-view)
        buf = np.zeros(rows * cols + 16, np.uint16)
        off = (-(buf.ctypes.data // 2) % 8 + misalign) % 8 if misalign else 0
        g = buf[off : off + rows * cols].reshape(rows, cols) if misalign else np.ascontiguousarray(grid, np.uint16).copy()
        g[...] = grid
        d = np.zeros((rows, cols), np.uint32)
        mx, st = ctypes.c_uint64(), ctypes.c_uint64()
        stats = np.zeros(8, np.uint64)
        rm, cm = rows // 2, cols // 2
        sr, sc = start or (rm + 1 - (rm % 2), cm + 1 - (cm % 2))
        rc = self.lib.emu_distance(rows, cols, self._p(g), sr, sc, self._p(d), ctypes.byref(mx), ctypes.byref(st), nwarps, sched, self._p(stats))
        assert rc == 0, f"emulated search failed: {rc}"
        return g, d, mx.value, st.value, stats

    def solver(self, game, grid, starts, finishes, nwarps=4, sched=0, want_paths=True):
        rows, cols = grid.shape
        g = np.ascontiguousarray(grid, np.uint16).copy()
        mp = rows * cols if want_paths else 0
        st = np.zeros((4, 2), np.int32)
        s = np.asarray(starts, np.int32).reshape(-1, 2)
        st[: len(s)] = s
        fi = np.zeros((4, 2), np.int32)
        f = np.asarray(finishes, np.int32).reshape(-1, 2)
        fi[: len(f)] = f
        win = ctypes.c_int32()
        claimed = np.zeros(4, np.int32)
        paths = np.zeros((4, max(mp, 1), 2), np.int32)
        lens = np.zeros(4, np.int64)
        painted = ctypes.c_uint64()
        stats = np.zeros(8, np.uint64)
        rc = self.lib.emu_solver(game, rows, cols, self._p(g), self._p(st), self._p(fi), ctypes.byref(win), self._p(claimed),
                                 self._p(paths), ctypes.c_longlong(mp), self._p(lens), ctypes.byref(painted), nwarps, sched, self._p(stats))
        assert rc == 0, f"emulated search failed: {rc}"
        return g, win.value, claimed, [paths[k, : lens[k] if want_paths else 0].copy() for k in range(4)], painted.value


def emu_band_engine(squares: np.ndarray, nwarps: int = 4, sched: int = 0):
    """The real band driver's engine interface (multithreading-with-mazes_b200/bands.py) over the emulator's
    emu_band_* entry points: the product's kernel bodies, host memory instead of device memory."""
    import torch

    sys.path.insert(0, ROOT)
    import __graft_entry__ as graft

    graft.load_package()
    from mazes_b200 import bands

    class EmuBandEngine(bands._BandEngine):
        prefix = "emu_band_"
        supports_lattice = True

        def grid_cols(self):
            return self.cols

        def __init__(self):
            self.torch = torch
            self.sq = np.ascontiguousarray(squares, np.uint16)
            self.rows, self.cols = self.sq.shape
            self.L = Emu().lib
            self.L.emu_band_create.restype = ctypes.c_void_p
            self.L.emu_band_levels.restype = ctypes.c_void_p
            self.h = ctypes.c_void_p(self.L.emu_band_create(self.rows, self.cols, self.sq.ctypes.data_as(ctypes.c_void_p), nwarps, sched))
            self.row_device = torch.device("cpu")
            self._setup_rows()

        def levels(self):
            ptr = self.L.emu_band_levels(self.h)
            return np.ctypeslib.as_array(ctypes.cast(ptr, ctypes.POINTER(ctypes.c_uint32)), shape=(self.rows, self.cols)).copy()

        def or_bits(self, cells, bits):
            for (r, c), b in zip(cells, bits):
                self.sq[r, c] |= np.uint16(b)

        def tree_used(self):
            return int(self.L.emu_band_tree_used(self.h))

        def close(self):
            self.L.emu_band_destroy(self.h)

    return EmuBandEngine()
