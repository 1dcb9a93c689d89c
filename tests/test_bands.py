"""Row-band sharding (N > 1).

* world_size-2 and -3 gloo runs of the REAL driver (multithreading-with-mazes_b200/bands.py) on the
  CPU with a numpy stand-in for the per-band search, checked against the oracle: this covers the
  host-side protocol (row exchange, border crossing, re-activation, termination).
* on a GPU: the lab_band_* entry points themselves, two and three bands driven in one process on
  one device (no rank ever waits on another inside a kernel), bit-exact against the oracle.
"""
import os
import sys
from collections import deque

import numpy as np
import pytest


def _leave():
    """End a spawned gloo worker without libtorch's exit handlers: under load they now and then crash in
    c10::Dispatcher::cleanup after the worker's results are already on disk (SIGSEGV in free, seen 2 of 6 runs)."""
    sys.stdout.flush()
    sys.stderr.flush()
    os._exit(0)


def _free_port():
    """A TCP port nobody holds right now (fixed ports collide with a finished test's TIME_WAIT sockets)."""
    import socket
    with socket.socket(socket.AF_INET, socket.SOCK_STREAM) as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
INF = 0xFFFFFFFF


class NumpyBandEngine:
    """Test double for GpuBandEngine: same interface, the band search is a plain label-correcting BFS."""

    def __init__(self, squares: np.ndarray):
        import torch

        self.torch = torch
        self.sq = squares
        self.rows, self.cols = squares.shape
        self.cols_padded = ((self.cols + 63) // 64) * 64
        self.row_device = torch.device("cpu")

    def begin(self, src_local):
        self.path = ((self.sq & 0x2000) != 0) & ((self.sq & 0x0200) == 0)
        self.dist = np.full((self.rows, self.cols), INF, np.uint64)
        self.ghost_up = np.full(self.cols, INF, np.uint64)
        self.ghost_dn = np.full(self.cols, INF, np.uint64)
        self.todo = deque()
        if src_local is not None:
            self.path[src_local] = True
            self.dist[src_local] = 0
            self.todo.append(src_local)

    def _row(self, a, dtype):
        out = np.zeros(self.cols_padded, dtype)
        out[: self.cols] = a
        return self.torch.from_numpy(out)

    def path_rows(self):
        return self._row(self.path[0].astype(np.uint8), np.uint8), self._row(self.path[-1].astype(np.uint8), np.uint8)

    def set_neighbour_paths(self, above_bottom, below_top):
        self.up_ok = above_bottom.numpy()[: self.cols].astype(bool) if above_bottom is not None else np.zeros(self.cols, bool)
        self.dn_ok = below_top.numpy()[: self.cols].astype(bool) if below_top is not None else np.zeros(self.cols, bool)

    def relax(self):
        d, p = self.dist, self.path
        while self.todo:
            r, c = self.todo.popleft()
            for nr, nc in ((r - 1, c), (r, c + 1), (r + 1, c), (r, c - 1)):
                if 0 <= nr < self.rows and 0 <= nc < self.cols and p[nr, nc] and d[nr, nc] > d[r, c] + 1:
                    d[nr, nc] = d[r, c] + 1
                    self.todo.append((nr, nc))

    def export_edges(self):
        return self._row(np.minimum(self.dist[0], INF).astype(np.int64).astype(np.uint32).view(np.int32), np.int32), \
            self._row(np.minimum(self.dist[-1], INF).astype(np.int64).astype(np.uint32).view(np.int32), np.int32)

    def import_edges(self, above_bottom, below_top):
        n = 0
        for src, row, ok in ((above_bottom, 0, self.up_ok), (below_top, self.rows - 1, self.dn_ok)):
            if src is None:
                continue
            v = src.numpy().view(np.uint32)[: self.cols].astype(np.uint64)
            for c in np.nonzero(ok & self.path[row] & (v != INF) & (v + 1 < self.dist[row]))[0]:
                self.dist[row, c] = v[c] + 1
                self.todo.append((row, int(c)))
                n += 1
        return n

    def finish(self):
        reached = self.dist != INF
        self.sq[reached] |= 0x0200
        return int(self.dist[reached].max()) if reached.any() else 0, int(reached.sum())

    def new_row(self, like):
        return self.torch.empty_like(like)


def _gloo_worker(rank, world, port, maze_path, out_dir):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist

    import __graft_entry__ as graft

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lab = graft.load_package()
    from mazes_b200 import bands

    maze = np.load(maze_path)
    r0, n = bands.split_rows(maze.shape[0], world)[rank]
    eng = NumpyBandEngine(maze[r0 : r0 + n].copy())
    res = bands.sharded_distance_map(eng, lab.distance_start(*maze.shape), r0, n)
    np.savez(os.path.join(out_dir, f"band{rank}.npz"), dist=np.minimum(eng.dist, INF).astype(np.uint32), sq=eng.sq, r0=r0,
             mx=res["max"], settled=res["settled"], rounds=res["rounds"])
    dist.destroy_process_group()
    _leave()


@pytest.mark.parametrize("builder,rows,cols,world", [("kruskal", 255, 131, 2), ("wilson", 193, 197, 3), ("grid", 257, 129, 2)])
def test_sharded_driver_on_gloo(lab, O, tmp_path, builder, rows, cols, world):
    import torch.multiprocessing as mp

    maze = lab.build_maze(builder, rows, cols, seed=31)
    path = str(tmp_path / "maze.npy")
    np.save(path, maze)
    port = _free_port()
    mp.spawn(_gloo_worker, args=(world, port, path, str(tmp_path)), nprocs=world, join=True)
    g_ref, d_ref, mx_ref, settled_ref = O.distance_map(maze)
    for rank in range(world):
        z = np.load(tmp_path / f"band{rank}.npz")
        r0 = int(z["r0"])
        n = z["dist"].shape[0]
        assert (z["dist"] == d_ref[r0 : r0 + n]).all() and (z["sq"] == g_ref[r0 : r0 + n]).all()
        assert int(z["mx"]) == mx_ref and int(z["settled"]) == settled_ref and int(z["rounds"]) >= 2


def _gloo_emu_worker(rank, world, port, maze_path, out_dir, use_tree):
    """Like _gloo_worker, but the band engine runs the PRODUCT's kernel bodies (tests/simt emulator): the wave's
    and -- for mazes that are one tree -- the sharded spanning-tree pipeline's."""
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist

    import __graft_entry__ as graft
    from _util import emu_band_engine

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lab = graft.load_package()
    from mazes_b200 import bands

    maze = np.load(maze_path)
    r0, n = bands.split_rows(maze.shape[0], world)[rank]
    eng = emu_band_engine(maze[r0 : r0 + n].copy(), nwarps=3, sched=rank + 1)
    res = bands.sharded_distance_map(eng, lab.distance_start(*maze.shape), r0, n, use_tree=use_tree)
    np.savez(os.path.join(out_dir, f"band{rank}.npz"), dist=eng.levels(), sq=eng.sq, r0=r0, mx=res["max"], settled=res["settled"],
             rounds=res["rounds"], road=res["road"], used=eng.tree_used())
    eng.close()
    dist.destroy_process_group()
    _leave()


def _gloo_emu_lattice_worker(rank, world, port, maze_path, out_dir, start):
    """The lattice road across bands of whole super-tile rows (bands.split_rows_super), kernel bodies on the emulator."""
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist

    import __graft_entry__ as graft
    from _util import emu_band_engine

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lab = graft.load_package()
    from mazes_b200 import bands

    maze = np.load(maze_path)
    r0, n = bands.split_rows_super(maze.shape[0], world)[rank]
    eng = emu_band_engine(maze[r0 : r0 + n].copy(), nwarps=3, sched=rank + 1)
    res = bands.sharded_distance_map(eng, tuple(start) if start else lab.distance_start(*maze.shape), r0, n)
    np.savez(os.path.join(out_dir, f"band{rank}.npz"), dist=eng.levels(), sq=eng.sq, r0=r0, mx=res["max"], settled=res["settled"],
             rounds=res["rounds"], road=res["road"], used=eng.tree_used())
    eng.close()
    dist.destroy_process_group()
    _leave()


def _check_bands(tmp_path, world, maze, O, road):
    g_ref, d_ref, mx_ref, settled_ref = O.distance_map(maze)
    for rank in range(world):
        z = np.load(tmp_path / f"band{rank}.npz")
        r0 = int(z["r0"])
        n = z["dist"].shape[0]
        assert str(z["road"]) == road and int(z["used"]) == {"wave": 0, "tree": 1, "room": 2, "lattice": 3}[road], (rank, str(z["road"]), int(z["used"]))
        assert (int(z["rounds"]) == 0) == (road != "wave")
        assert (z["dist"] == d_ref[r0 : r0 + n]).all(), (rank, int((z["dist"] != d_ref[r0 : r0 + n]).sum()))
        assert (z["sq"] == g_ref[r0 : r0 + n]).all()
        assert int(z["mx"]) == mx_ref and int(z["settled"]) == settled_ref


@pytest.mark.parametrize("builder,rows,cols,world,start", [("kruskal", 511, 131, 2, None), ("wilson", 767, 67, 3, None), ("rdfs", 513, 129, 2, (511, 1)),
                                                          ("kruskal", 767, 65, 3, (1, 63))])
def test_sharded_lattice_road_on_gloo(lab, O, tmp_path, builder, rows, cols, world, start):
    """Reference builders' mazes across 2-3 bands of whole super-tile rows: the lattice road, two exchanges per map,
    bit-exact.  (513 rows over 2 ranks: the last band is ONE row; starts in the first / last band, next to a band border.)"""
    import torch.multiprocessing as mp

    maze = lab.build_maze(builder, rows, cols, seed=37)
    path = str(tmp_path / "maze.npy")
    np.save(path, maze)
    port = _free_port()
    mp.spawn(_gloo_emu_lattice_worker, args=(world, port, path, str(tmp_path), start), nprocs=world, join=True)
    if start is None:
        _check_bands(tmp_path, world, maze, O, "lattice")
    else:
        d_ref = O.bfs_levels(maze, *start)
        for rank in range(world):
            z = np.load(tmp_path / f"band{rank}.npz")
            r0, n = int(z["r0"]), z["dist"].shape[0]
            assert str(z["road"]) == "lattice" and int(z["used"]) == 3
            assert (z["dist"] == d_ref[r0 : r0 + n]).all(), (rank, int((z["dist"] != d_ref[r0 : r0 + n]).sum()))


def test_sharded_lattice_that_is_no_tree_falls_back_on_gloo(lab, O, tmp_path):
    import torch.multiprocessing as mp

    from test_tree_pipeline import tree_looking_non_tree

    maze = tree_looking_non_tree(lab, 511, 67, 2)
    path = str(tmp_path / "maze.npy")
    np.save(path, maze)
    port = _free_port()
    mp.spawn(_gloo_emu_lattice_worker, args=(2, port, path, str(tmp_path), None), nprocs=2, join=True)
    _check_bands(tmp_path, 2, maze, O, "wave")


@pytest.mark.parametrize("builder,rows,cols,world", [("kruskal", 255, 131, 2), ("wilson", 193, 197, 3), ("rdfs", 257, 67, 2), ("kruskal", 447, 65, 3)])
def test_sharded_tree_pipeline_on_gloo(lab, O, tmp_path, builder, rows, cols, world):
    """One tree across 2-3 bands: the sharded spanning-tree pipeline (no relax/exchange round at all), bit-exact.
    (447 rows over 3 ranks: uneven bands; the start lies in the first, middle or last band.)"""
    import torch.multiprocessing as mp

    maze = lab.build_maze(builder, rows, cols, seed=31)
    path = str(tmp_path / "maze.npy")
    np.save(path, maze)
    port = _free_port()
    mp.spawn(_gloo_emu_worker, args=(world, port, path, str(tmp_path), True), nprocs=world, join=True)
    _check_bands(tmp_path, world, maze, O, "tree")


def test_sharded_non_trees_fall_back_to_the_wave_on_gloo(lab, O, tmp_path):
    """Not one tree: the counts say so (grid + cross), or only the later checks do (a cycle plus a cut-off part with
    a tree's counts): every rank gives the field back and the relax/exchange loop does the search."""
    import torch.multiprocessing as mp

    from test_tree_pipeline import tree_looking_non_tree

    for i, maze in enumerate([lab.build_maze("grid", 257, 129, seed=31, mod="cross"), tree_looking_non_tree(lab, 255, 131, 2)]):
        path = str(tmp_path / "maze.npy")
        np.save(path, maze)
        port = _free_port()
        mp.spawn(_gloo_emu_worker, args=(2, port, path, str(tmp_path), True), nprocs=2, join=True)
        _check_bands(tmp_path, 2, maze, O, "wave")


def test_sharded_open_room_on_gloo(lab, O, tmp_path):
    """An arena across 3 bands: no search and no exchange, every band writes its part of the room from the closed form
    (the start lies in the middle band; for the other two it is above / below the band).  One wall square inside
    and it is the relax/exchange loop again."""
    import torch.multiprocessing as mp

    arena = lab.build_maze("arena", 193, 131, seed=1)
    holed = arena.copy()
    holed[150, 20] &= ~np.uint16(lab.PATH_BIT)
    for i, (maze, road) in enumerate([(arena, "room"), (holed, "wave")]):
        path = str(tmp_path / "maze.npy")
        np.save(path, maze)
        port = _free_port()
        mp.spawn(_gloo_emu_worker, args=(3, port, path, str(tmp_path), True), nprocs=3, join=True)
        _check_bands(tmp_path, 3, maze, O, road)


def _gloo_game_worker(rank, world, port, maze_path, out_dir, game, spec):
    """A sharded game (bands.sharded_hunt / sharded_gather) with the product's kernel bodies on the emulator."""
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist

    import __graft_entry__ as graft
    from _util import emu_band_engine

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    graft.load_package()
    from mazes_b200 import bands

    maze = np.load(maze_path)
    r0, n = (bands.split_rows_super if spec.get("super") else bands.split_rows)(maze.shape[0], world)[rank]
    eng = emu_band_engine(maze[r0 : r0 + n].copy(), nwarps=3, sched=rank + 1)
    if game == "hunt":
        res = bands.sharded_hunt(eng, spec["start"], spec["finish"], r0, n)
        extra = dict(winner=res["winner"], path=res["path"], path_len=res["path_len"])
    elif game == "corners":
        res = bands.sharded_corners(eng, spec["starts"], spec["finish"], r0, n)
        extra = dict(winner=res["winner"], level=res["level"], rounds=res["rounds"])
    else:
        res = bands.sharded_gather(eng, spec["start"], spec["finishes"], r0, n)
        extra = dict(claimed_by=np.asarray(res["claimed_by"]), path_len=np.asarray(res["path_len"]),
                     **{f"path{k}": (res["paths"][k] if res["paths"][k] is not None else np.zeros((0, 2), np.int32)) for k in range(4)})
    np.savez(os.path.join(out_dir, f"band{rank}.npz"), sq=eng.sq, r0=r0, road=res["road"], painted=res["painted"], **extra)
    eng.close()
    dist.destroy_process_group()
    _leave()


@pytest.mark.parametrize("builder,rows,cols,world,road", [("kruskal", 255, 131, 2, "tree"), ("wilson", 193, 197, 3, "tree"), ("arena", 193, 131, 3, "room"),
                                                          ("rdfs", 447, 67, 3, "tree")])  # (rdfs: corridors that cross the band borders again and again)
def test_sharded_hunt_and_gather_on_gloo(lab, O, tmp_path, builder, rows, cols, world, road):
    """Bfs::hunt and Bfs::gather across 2-3 bands (one tree / an open room): grids, winner, claims and the canonical paths
    -- walked band by band, handed over at the borders -- equal the single-maze canonical model's."""
    import torch.multiprocessing as mp

    maze = lab.build_maze(builder, rows, cols, seed=41)
    for i, seed in enumerate((7,)):
        # hunt
        placed, s, f = lab.place_hunt(maze, seed=seed)
        path = str(tmp_path / "maze.npy")
        np.save(path, placed)
        spec = {"start": tuple(int(x) for x in s), "finish": tuple(int(x) for x in f)}
        mp.spawn(_gloo_game_worker, args=(world, _free_port(), path, str(tmp_path), "hunt", spec), nprocs=world, join=True)
        g_ref, w_ref, p_ref = O.canon_hunt(placed, s, f)
        z = [np.load(tmp_path / f"band{r}.npz") for r in range(world)]
        assert all(str(b["road"]) == road for b in z)
        assert (np.concatenate([b["sq"] for b in z]) == g_ref).all()
        for b in z:
            assert int(b["winner"]) == w_ref and (b["path"] == np.asarray(p_ref)).all() and int(b["path_len"]) == len(p_ref)
        assert int(z[0]["painted"]) == int(((g_ref & 0x00F0) != 0).sum())
        # gather
        placed, s, fin = lab.place_gather(maze, seed=seed + 20)
        np.save(path, placed)
        spec = {"start": tuple(int(x) for x in s), "finishes": [tuple(int(x) for x in q) for q in fin]}
        mp.spawn(_gloo_game_worker, args=(world, _free_port(), path, str(tmp_path), "gather", spec), nprocs=world, join=True)
        g_ref, c_ref, p_ref = O.canon_gather(placed, s, fin)
        z = [np.load(tmp_path / f"band{r}.npz") for r in range(world)]
        got = np.concatenate([b["sq"] for b in z])
        assert (got == g_ref).all(), int((got != g_ref).sum())
        for b in z:
            assert (b["claimed_by"] == c_ref).all()
            for k in range(4):
                assert (b[f"path{k}"] == np.asarray(p_ref[k]).reshape(-1, 2)).all()


@pytest.mark.parametrize("builder,rows,cols,world", [("kruskal", 767, 131, 3), ("wilson", 513, 195, 2), ("rdfs", 1023, 67, 2)])
def test_sharded_hunt_on_the_lattice_road_on_gloo(lab, O, tmp_path, builder, rows, cols, world):
    """Bfs::hunt across bands of whole super-tile rows: the lattice road's level field, and the winner's path WITHOUT any
    hand-over between the bands -- every band marks the segments in its own rows from the whole-maze tour ranks
    (lab_band_lat_path_rank / _mark).  Grid, winner and the canonical path equal the single-maze model's."""
    import torch.multiprocessing as mp

    maze = lab.build_maze(builder, rows, cols, seed=59)
    done = 0
    for seed in range(40):
        placed, s, f = lab.place_hunt(maze, seed=seed)
        if not (s[0] % 2 == 1 and s[1] % 2 == 1):
            continue  # (the lattice road across bands wants the start on a cell; other starts take the general pipeline)
        path = str(tmp_path / "maze.npy")
        np.save(path, placed)
        spec = {"start": tuple(int(x) for x in s), "finish": tuple(int(x) for x in f), "super": True}
        mp.spawn(_gloo_game_worker, args=(world, _free_port(), path, str(tmp_path), "hunt", spec), nprocs=world, join=True)
        g_ref, w_ref, p_ref = O.canon_hunt(placed, s, f)
        z = [np.load(tmp_path / f"band{r}.npz") for r in range(world)]
        assert all(str(b["road"]) == "lattice" for b in z)
        got = np.concatenate([b["sq"] for b in z])
        assert (got == g_ref).all(), int((got != g_ref).sum())
        for b in z:
            assert int(b["winner"]) == w_ref and int(b["path_len"]) == len(p_ref) and (b["path"] == np.asarray(p_ref)).all()
        done += 1
        if done == 2:
            break
    assert done == 2


def test_sharded_gather_on_the_lattice_road_on_gloo(lab, O, tmp_path):
    """Bfs::gather across bands of whole super-tile rows: the lattice road's field, claims and path.front() recolours by the
    owning bands, the four paths walked band by band on that field."""
    import torch.multiprocessing as mp

    maze = lab.build_maze("kruskal", 767, 131, seed=61)
    seed = next(q for q in range(100) if all(int(x) % 2 == 1 for x in lab.place_gather(maze, seed=q)[1]))
    placed, s, fin = lab.place_gather(maze, seed=seed)
    path = str(tmp_path / "maze.npy")
    np.save(path, placed)
    spec = {"start": tuple(int(x) for x in s), "finishes": [tuple(int(x) for x in q) for q in fin], "super": True}
    mp.spawn(_gloo_game_worker, args=(3, _free_port(), path, str(tmp_path), "gather", spec), nprocs=3, join=True)
    g_ref, c_ref, p_ref = O.canon_gather(placed, s, fin)
    z = [np.load(tmp_path / f"band{r}.npz") for r in range(3)]
    assert all(str(b["road"]) == "lattice" for b in z)
    got = np.concatenate([b["sq"] for b in z])
    assert (got == g_ref).all(), int((got != g_ref).sum())
    for b in z:
        assert (b["claimed_by"] == c_ref).all()
        for k in range(4):
            assert (b[f"path{k}"] == np.asarray(p_ref[k]).reshape(-1, 2)).all()


def _hunt_and_gather_checks(lab, O, tmp_path, maze, world, road, seed=7):
    import torch.multiprocessing as mp

    placed, s, f = lab.place_hunt(maze, seed=seed)
    path = str(tmp_path / "maze.npy")
    np.save(path, placed)
    spec = {"start": tuple(int(x) for x in s), "finish": tuple(int(x) for x in f)}
    mp.spawn(_gloo_game_worker, args=(world, _free_port(), path, str(tmp_path), "hunt", spec), nprocs=world, join=True)
    g_ref, w_ref, p_ref = O.canon_hunt(placed, s, f)
    z = [np.load(tmp_path / f"band{r}.npz") for r in range(world)]
    assert all(str(b["road"]) == road for b in z)
    assert (np.concatenate([b["sq"] for b in z]) == g_ref).all()
    for b in z:
        assert int(b["winner"]) == w_ref and (b["path"] == np.asarray(p_ref)).all() and int(b["path_len"]) == len(p_ref)
    placed, s, fin = lab.place_gather(maze, seed=seed + 20)
    np.save(path, placed)
    spec = {"start": tuple(int(x) for x in s), "finishes": [tuple(int(x) for x in q) for q in fin]}
    mp.spawn(_gloo_game_worker, args=(world, _free_port(), path, str(tmp_path), "gather", spec), nprocs=world, join=True)
    g_ref, c_ref, p_ref = O.canon_gather(placed, s, fin)
    z = [np.load(tmp_path / f"band{r}.npz") for r in range(world)]
    got = np.concatenate([b["sq"] for b in z])
    assert all(str(b["road"]) == road for b in z)
    assert (got == g_ref).all(), int((got != g_ref).sum())
    for b in z:
        assert (b["claimed_by"] == c_ref).all()
        for k in range(4):
            assert (b[f"path{k}"] == np.asarray(p_ref[k]).reshape(-1, 2)).all()


@pytest.mark.parametrize("builder,mod,rows,cols,world", [("grid", "cross", 191, 131, 2), ("kruskal", "x", 193, 133, 3)])
def test_sharded_hunt_and_gather_on_general_mazes_on_gloo(lab, O, tmp_path, builder, mod, rows, cols, world):
    """Mazes with loops (-m cross, -m x: bfs_threads.cc:35-85,144-185 on anything the builders make) across 2-3 bands: the
    field comes from the tile wave run band by band to its fixpoint, the games' thresholds are applied to it afterwards --
    grids, claims and canonical paths equal the single-maze canonical model's."""
    maze = lab.build_maze(builder, rows, cols, seed=43, mod=mod)
    _hunt_and_gather_checks(lab, O, tmp_path, maze, world, "wave")


@pytest.mark.parametrize("builder,mod,rows,cols,world", [("grid", "cross", 191, 131, 2), ("grid", "cross", 257, 67, 3), ("kruskal", None, 193, 131, 3),
                                                         ("arena", None, 129, 131, 2)])
def test_sharded_corners_on_gloo(lab, O, tmp_path, builder, mod, rows, cols, world):
    """Bfs::corners (bfs_threads.cc:420-453; BASELINE configs[3] is grid + cross) across 2-3 bands: four fields searched at
    once by the wave, four rows of levels per exchange; grid and winner equal the single-maze canonical model's."""
    import torch.multiprocessing as mp

    maze = lab.build_maze(builder, rows, cols, seed=47, mod=mod)
    placed, starts, finish = lab.place_corners(maze, seed=5)
    path = str(tmp_path / "maze.npy")
    np.save(path, placed)
    spec = {"starts": [tuple(int(x) for x in q) for q in starts], "finish": tuple(int(x) for x in finish)}
    mp.spawn(_gloo_game_worker, args=(world, _free_port(), path, str(tmp_path), "corners", spec), nprocs=world, join=True)
    g_ref, w_ref, p_ref = O.canon_corners(placed, starts, finish)
    z = [np.load(tmp_path / f"band{r}.npz") for r in range(world)]
    got = np.concatenate([b["sq"] for b in z])
    assert (got == g_ref).all(), int((got != g_ref).sum())
    for b in z:
        assert int(b["winner"]) == w_ref and int(b["level"]) == len(p_ref)
    assert int(z[0]["painted"]) == int(((g_ref & 0x00F0) != 0).sum())


def test_split_rows(lab):
    from mazes_b200 import bands

    assert bands.split_rows(65535, 8) == [(i * 8192, 8192) for i in range(7)] + [(57344, 8191)]
    assert bands.split_rows(255, 2) == [(0, 128), (128, 127)]
    with pytest.raises(ValueError):
        bands.split_rows(63, 2)


@pytest.mark.gpu
@pytest.mark.parametrize("builder,rows,cols,nb", [("kruskal", 511, 383, 2), ("wilson", 767, 257, 3), ("arena", 255, 255, 2), ("grid", 641, 513, 4)])
def test_band_entry_points_one_process(lab, O, builder, rows, cols, nb):
    """The lab_band_* kernels, several bands on ONE device driven in lock step by this process."""
    import torch

    from mazes_b200 import bands

    maze = lab.build_maze(builder, rows, cols, seed=77, mod="cross" if builder == "grid" else None)
    split = bands.split_rows(rows, nb)
    sq = [torch.from_numpy(maze[r0 : r0 + n].view(np.int16).copy()).cuda() for r0, n in split]
    lv = [torch.empty((n, cols), dtype=torch.int32, device="cuda") for _, n in split]
    eng = [bands.GpuBandEngine(lab, s) for s in sq]
    for e, l in zip(eng, lv):
        e.grid.bind_levels(0, l.data_ptr())
    sr, sc = lab.distance_start(rows, cols)
    for e, (r0, n) in zip(eng, split):
        e.begin((sr - r0, sc) if r0 <= sr < r0 + n else None)
    rows_p = [tuple(t.clone() for t in e.path_rows()) for e in eng]
    for i, e in enumerate(eng):
        e.set_neighbour_paths(rows_p[i - 1][1] if i > 0 else None, rows_p[i + 1][0] if i < nb - 1 else None)
    rounds = 0
    while True:
        rounds += 1
        for e in eng:
            e.relax()
        edges = [tuple(t.clone() for t in e.export_edges()) for e in eng]
        act = [e.import_edges(edges[i - 1][1] if i > 0 else None, edges[i + 1][0] if i < nb - 1 else None) for i, e in enumerate(eng)]
        if not any(act):
            break
        assert rounds < 100000
    res = [e.finish() for e in eng]
    g_ref, d_ref, mx_ref, settled_ref = O.distance_map(maze)
    got_d = np.concatenate([l.cpu().numpy().view(np.uint32) for l in lv])
    got_g = np.concatenate([s.cpu().numpy().view(np.uint16) for s in sq])
    assert (got_d == d_ref).all() and (got_g == g_ref).all()
    assert max(r[0] for r in res) == mx_ref and sum(r[1] for r in res) == settled_ref


@pytest.mark.gpu
@pytest.mark.parametrize("builder,mod,rows,cols,nb", [("grid", "cross", 511, 387, 2), ("grid", "cross", 1023, 643, 4), ("kruskal", "x", 447, 259, 3),
                                                     ("arena", None, 383, 321, 3)])
def test_gpu_corners_across_bands_one_process(lab, O, builder, mod, rows, cols, nb):
    """Bfs::corners across 2-4 bands on the CUDA band engine (lab_band_begin_corners, four rows of levels per exchange),
    several bands driven in lock step by one process: grid and winner equal the canonical model's."""
    import torch

    from mazes_b200 import bands

    maze = lab.build_maze(builder, rows, cols, seed=53, mod=mod)
    placed, starts, finish = lab.place_corners(maze, seed=9)
    split = bands.split_rows(rows, nb)
    sq = [torch.from_numpy(placed[r0 : r0 + n].view(np.int16).copy()).cuda() for r0, n in split]
    eng = [bands.GpuBandEngine(lab, s) for s in sq]
    for e, (r0, n) in zip(eng, split):
        e.begin_corners([(int(r) - r0, int(c)) if r0 <= r < r0 + n else None for r, c in starts])
        assert e.L.lab_band_planes(e.h) == 4
    rows_p = [tuple(t.clone() for t in e.path_rows()) for e in eng]
    for i, e in enumerate(eng):
        e.set_neighbour_paths(rows_p[i - 1][1] if i > 0 else None, rows_p[i + 1][0] if i < nb - 1 else None)
    rounds = 0
    while True:
        rounds += 1
        for e in eng:
            e.relax()
        edges = [tuple(t.clone() for t in e.export_edges()) for e in eng]
        act = [e.import_edges(edges[i - 1][1] if i > 0 else None, edges[i + 1][0] if i < nb - 1 else None) for i, e in enumerate(eng)]
        if not any(act):
            break
        assert rounds < 100000
    fr, fc = int(finish[0]), int(finish[1])
    owner = next(i for i, (r0, n) in enumerate(split) if r0 <= fr < r0 + n)
    levels = [eng[owner].level_at(fr - split[owner][0], fc, k) for k in range(4)]
    level = min(levels)
    winner = levels.index(level) if level != bands.INF else -1
    painted = sum(e.paint([(level, (1 << k) << bands.PAINT_SHIFT, k) for k in range(4)]) for e in eng)
    for e in eng:
        e.finish_game()
    g_ref, w_ref, p_ref = O.canon_corners(placed, starts, finish)
    got = np.concatenate([s.cpu().numpy().view(np.uint16) for s in sq])
    assert (got == g_ref).all(), int((got != g_ref).sum())
    assert winner == w_ref and level == len(p_ref) and painted == int(((g_ref & 0x00F0) != 0).sum())
    # ... and equal to the one-GPU game on the whole maze
    with lab.Grid.from_host(placed) as g:
        res = g.corners(starts, finish, want_path=False)
        assert (g.download() == g_ref).all() and res.winner == w_ref


def _lockstep_tree(eng, split, start, cols):
    """_sharded_tree of bands.py for several bands driven by ONE process: an all-gather becomes the concatenation
    of the bands' chunks, copied into every band's receive buffer."""
    import torch

    from mazes_b200 import bands

    nb = len(eng)
    lay = bands.tree_layout(split, eng[0].cols_padded // 64)
    assert lay is not None
    bufs = [e.tree_bind(lay, i) for i, e in enumerate(eng)]

    def all_gather(stage):
        every = torch.cat([b[f"chunk{stage}"] for b in bufs])
        for b in bufs:
            b[f"all{stage}"].copy_(every)

    counts = [e.tree_count() for e in eng]
    squares, pairs = sum(c[0] for c in counts), sum(c[1] for c in counts)
    if squares < 2 or pairs + 1 != squares:
        boxes = [(r0 + c[3], r0 + c[4], c[5], c[6]) for c, (r0, _) in zip(counts, split) if c[0]]
        R0, R1, C0, C1 = min(b[0] for b in boxes), max(b[1] for b in boxes), min(b[2] for b in boxes), max(b[3] for b in boxes)
        if squares != (R1 - R0 + 1) * (C1 - C0 + 1):
            return False
        for e, (r0, n) in zip(eng, split):
            e.room(max(R0, r0) - r0, min(R1, r0 + n - 1) - r0, C0, C1, start[0] - r0, start[1])
        return "room"
    owner = next(i for i, (r0, n) in enumerate(split) if r0 <= start[0] < r0 + n)
    for e, b in zip(eng, bufs):
        e.tree_rank_local(squares, pairs, b["chunk1"])
    all_gather(1)
    for e, b in zip(eng, bufs):
        e.tree_flood(b["all1"], owner, b["chunk2"])
    all_gather(2)
    for e, b in zip(eng, bufs):
        e.tree_prefix_local(b["all2"], b["chunk3"])
    all_gather(3)
    used = [e.tree_levels(b["all3"]) for e, b in zip(eng, bufs)]
    assert all(u == used[0] for u in used)
    return used[0]


@pytest.mark.gpu
@pytest.mark.parametrize("builder,rows,cols,nb", [("kruskal", 511, 383, 2), ("wilson", 767, 257, 3), ("rdfs", 1023, 515, 4), ("kruskal", 2047, 1025, 8)])
def test_band_tree_entry_points_one_process(lab, O, builder, rows, cols, nb):
    """The lab_band_tree_* kernels: one tree across 2-8 bands on ONE device, driven in lock step by this process."""
    import torch

    from mazes_b200 import bands

    maze = lab.build_maze(builder, rows, cols, seed=78)
    split = bands.split_rows(rows, nb)
    sq = [torch.from_numpy(maze[r0 : r0 + n].view(np.int16).copy()).cuda() for r0, n in split]
    lv = [torch.full((n, cols), 0x5A5A5A5A, dtype=torch.int32, device="cuda") for _, n in split]
    eng = [bands.GpuBandEngine(lab, s) for s in sq]
    for e, l in zip(eng, lv):
        e.grid.bind_levels(0, l.data_ptr())
    start = lab.distance_start(rows, cols)
    for e, (r0, n) in zip(eng, split):
        e.begin((start[0] - r0, start[1]) if r0 <= start[0] < r0 + n else None)
    rows_p = [tuple(t.clone() for t in e.path_rows()) for e in eng]
    for i, e in enumerate(eng):
        e.set_neighbour_paths(rows_p[i - 1][1] if i > 0 else None, rows_p[i + 1][0] if i < nb - 1 else None)
    assert _lockstep_tree(eng, split, start, cols)
    res = [e.finish() for e in eng]
    assert all(e.grid.stats()["tree_used"] == 1 for e in eng)
    g_ref, d_ref, mx_ref, settled_ref = O.distance_map(maze)
    got_d = np.concatenate([l.cpu().numpy().view(np.uint32) for l in lv])
    got_g = np.concatenate([s.cpu().numpy().view(np.uint16) for s in sq])
    assert (got_d == d_ref).all() and (got_g == g_ref).all()
    assert max(r[0] for r in res) == mx_ref and sum(r[1] for r in res) == settled_ref


@pytest.mark.gpu
def test_band_tree_gives_a_non_tree_back_to_the_wave(lab, O):
    """A maze with a tree's counts that is not one: the flood says so, every band wipes its field and the relax /
    exchange loop of test_band_entry_points_one_process takes over."""
    import torch

    from mazes_b200 import bands
    from test_tree_pipeline import tree_looking_non_tree

    rows, cols, nb = 511, 767, 3
    maze = tree_looking_non_tree(lab, rows, cols, 3)
    split = bands.split_rows(rows, nb)
    sq = [torch.from_numpy(maze[r0 : r0 + n].view(np.int16).copy()).cuda() for r0, n in split]
    lv = [torch.empty((n, cols), dtype=torch.int32, device="cuda") for _, n in split]
    eng = [bands.GpuBandEngine(lab, s) for s in sq]
    for e, l in zip(eng, lv):
        e.grid.bind_levels(0, l.data_ptr())
    start = lab.distance_start(rows, cols)
    for e, (r0, n) in zip(eng, split):
        e.begin((start[0] - r0, start[1]) if r0 <= start[0] < r0 + n else None)
    rows_p = [tuple(t.clone() for t in e.path_rows()) for e in eng]
    for i, e in enumerate(eng):
        e.set_neighbour_paths(rows_p[i - 1][1] if i > 0 else None, rows_p[i + 1][0] if i < nb - 1 else None)
    assert not _lockstep_tree(eng, split, start, cols)
    rounds = 0
    while True:
        rounds += 1
        for e in eng:
            e.relax()
        edges = [tuple(t.clone() for t in e.export_edges()) for e in eng]
        act = [e.import_edges(edges[i - 1][1] if i > 0 else None, edges[i + 1][0] if i < nb - 1 else None) for i, e in enumerate(eng)]
        if not any(act):
            break
        assert rounds < 100000
    res = [e.finish() for e in eng]
    g_ref, d_ref, mx_ref, settled_ref = O.distance_map(maze)
    got_d = np.concatenate([l.cpu().numpy().view(np.uint32) for l in lv])
    got_g = np.concatenate([s.cpu().numpy().view(np.uint16) for s in sq])
    assert (got_d == d_ref).all() and (got_g == g_ref).all()
    assert max(r[0] for r in res) == mx_ref and sum(r[1] for r in res) == settled_ref


@pytest.mark.gpu
@pytest.mark.parametrize("rows,cols,nb,start", [(511, 383, 3, None), (1023, 2049, 4, (1000, 7)), (129, 67, 2, (1, 65))])
def test_band_room_one_process(lab, O, rows, cols, nb, start):
    """lab_band_room: an arena across 2-4 bands on ONE device; every band writes its part of the room from the closed
    form, whichever band the start lies in."""
    import torch

    from mazes_b200 import bands

    maze = lab.build_maze("arena", rows, cols, seed=1)
    split = bands.split_rows(rows, nb)
    sq = [torch.from_numpy(maze[r0 : r0 + n].view(np.int16).copy()).cuda() for r0, n in split]
    lv = [torch.full((n, cols), 0x5A5A5A5A, dtype=torch.int32, device="cuda") for _, n in split]
    eng = [bands.GpuBandEngine(lab, s) for s in sq]
    for e, l in zip(eng, lv):
        e.grid.bind_levels(0, l.data_ptr())
    start = start or lab.distance_start(rows, cols)
    for e, (r0, n) in zip(eng, split):
        e.begin((start[0] - r0, start[1]) if r0 <= start[0] < r0 + n else None)
    rows_p = [tuple(t.clone() for t in e.path_rows()) for e in eng]
    for i, e in enumerate(eng):
        e.set_neighbour_paths(rows_p[i - 1][1] if i > 0 else None, rows_p[i + 1][0] if i < nb - 1 else None)
    assert _lockstep_tree(eng, split, start, cols) == "room"
    res = [e.finish() for e in eng]
    assert all(e.grid.stats()["tree_used"] == 2 for e in eng)
    d_ref = O.bfs_levels(maze, *start)
    reached = d_ref != INF
    got_d = np.concatenate([l.cpu().numpy().view(np.uint32) for l in lv])
    got_g = np.concatenate([s.cpu().numpy().view(np.uint16) for s in sq])
    assert (got_d == d_ref).all() and (((got_g & lab.MEASURE_BIT) != 0) == reached).all()
    assert max(r[0] for r in res) == int(d_ref[reached].max()) and sum(r[1] for r in res) == int(reached.sum())


@pytest.mark.gpu
def test_gpu_sharded_games_single_rank(lab, O):
    """bands.sharded_hunt / sharded_gather with the CUDA band engine as a one-rank job (the lab_band_begin_game, _level_at,
    _paint, _walk, _recolour entry points; hand-overs between ranks are covered by the gloo tests on the same kernel bodies
    and by tools/run_bands.py --game ... --check over NCCL)."""
    import torch
    import torch.distributed as dist

    from mazes_b200 import bands

    created = not dist.is_initialized()
    if created:
        os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(_free_port()))
        dist.init_process_group("gloo", rank=0, world_size=1)
    try:
        for builder, rows, cols in [("kruskal", 511, 767), ("arena", 383, 1025), ("rdfs", 257, 259)]:
            maze = lab.build_maze(builder, rows, cols, seed=5)
            placed, s, f = lab.place_hunt(maze, seed=11)
            sq = torch.from_numpy(placed.view(np.int16).copy()).cuda()
            eng = bands.GpuBandEngine(lab, sq)
            res = bands.sharded_hunt(eng, tuple(int(x) for x in s), tuple(int(x) for x in f), 0, rows)
            g_ref, w_ref, p_ref = O.canon_hunt(placed, s, f)
            assert (sq.cpu().numpy().view(np.uint16) == g_ref).all() and res["winner"] == w_ref and (res["path"] == np.asarray(p_ref)).all()
            if builder != "arena":  # a start on a cell: the lattice road, the winner's path from the tour's ranks (lab_band_lat_path_*)
                seed = next(q for q in range(100) if all(int(x) % 2 == 1 for x in lab.place_hunt(maze, seed=q)[1]))
                placed, s, f = lab.place_hunt(maze, seed=seed)
                sq = torch.from_numpy(placed.view(np.int16).copy()).cuda()
                eng = bands.GpuBandEngine(lab, sq)
                res = bands.sharded_hunt(eng, tuple(int(x) for x in s), tuple(int(x) for x in f), 0, rows)
                g_ref, w_ref, p_ref = O.canon_hunt(placed, s, f)
                assert res["road"] == "lattice"
                assert (sq.cpu().numpy().view(np.uint16) == g_ref).all() and res["winner"] == w_ref and (res["path"] == np.asarray(p_ref)).all()
            placed, s, fin = lab.place_gather(maze, seed=12)
            sq = torch.from_numpy(placed.view(np.int16).copy()).cuda()
            eng = bands.GpuBandEngine(lab, sq)
            res = bands.sharded_gather(eng, tuple(int(x) for x in s), [tuple(int(x) for x in q) for q in fin], 0, rows)
            g_ref, c_ref, p_ref = O.canon_gather(placed, s, fin)
            assert (sq.cpu().numpy().view(np.uint16) == g_ref).all() and (np.asarray(res["claimed_by"]) == c_ref).all()
            assert all((res["paths"][k] == np.asarray(p_ref[k]).reshape(-1, 2)).all() for k in range(4))
    finally:
        if created:
            dist.destroy_process_group()


@pytest.mark.parametrize("builder,rows,cols,world", [("kruskal", 511, 131, 2), ("wilson", 1023, 67, 4), ("rdfs", 767, 129, 3)])
def test_lattice_bands_peer_mode_one_process(lab, O, builder, rows, cols, world):
    """lab_band_lat_set_peers: every band's kernels store its slice of the border list and its weights straight into EVERY
    band's copy (what lab_distance_map_sharded does over NVLink between the devices of one process): no all-gather, no
    all-reduce, the phases only have to be ordered.  Kernel bodies on the emulator, bands in lock step."""
    import ctypes

    import torch

    from _util import emu_band_engine
    from mazes_b200 import bands

    maze = lab.build_maze(builder, rows, cols, seed=41)
    split = bands.split_rows_super(rows, world)
    start = lab.distance_start(rows, cols)
    engs = [emu_band_engine(maze[r0:r0 + n].copy(), nwarps=3, sched=r + 1) for r, (r0, n) in enumerate(split)]
    for e, (r0, n) in zip(engs, split):
        e.begin((start[0] - r0, start[1]) if r0 <= start[0] < r0 + n else None, lattice_next=True)
    rows_ = [e.path_rows() for e in engs]
    tops, bots = [r[0].clone() for r in rows_], [r[1].clone() for r in rows_]
    for r, e in enumerate(engs):
        e.set_neighbour_paths(bots[r - 1] if r > 0 else None, tops[r + 1] if r < world - 1 else None)
    per = (split[0][1] + 255) // 256
    bufs = [e.lat_bind(r, per, world) for r, e in enumerate(engs)]
    for r, e in enumerate(engs):
        order = [r] + [q for q in range(world) if q != r]
        gbs = (ctypes.c_void_p * world)(*[bufs[q]["gb"].data_ptr() for q in order])
        ws = (ctypes.c_void_p * world)(*[bufs[q]["wsum"].data_ptr() for q in order])
        assert e.L.emu_band_lat_set_peers(e.h, world, gbs, ws) == 0
        bufs[r]["wsum"].fill_(0x5A5A5A5A)  # (peer mode does not clear the weights: every position is written by its owner)
        tail = bufs[r]["reduce_words"] - 8
        bufs[r]["wsum"][tail:tail + 8] = 0
    for e, b in zip(engs, bufs):
        e.lat_link(b["hdr"])
    all_hdr = torch.cat([b["hdr"] for b in bufs])
    for e, b in zip(engs, bufs):
        b["all_hdr"].copy_(all_hdr)
    for e, b in zip(engs, bufs):
        e.lat_flood(b["all_hdr"], world, False)
    assert all(e.lat_levels() for e in engs)
    g_ref, d_ref, mx_ref, settled_ref = O.distance_map(maze)
    mx, settled = 0, 0
    for e, (r0, n) in zip(engs, split):
        m, s = e.finish()
        mx, settled = max(mx, m), settled + s
        assert (e.levels() == d_ref[r0:r0 + n]).all() and (e.sq == g_ref[r0:r0 + n]).all()
        e.close()
    assert mx == mx_ref and settled == settled_ref


@pytest.mark.gpu
def test_gpu_distance_map_sharded_one_process(lab, O):
    """lab_distance_map_sharded through the C ABI.  With one visible GPU both bands run on it (device list [0, 0]: the peer
    stores are then ordinary stores); with two or more, on two devices over NVLink."""
    import torch

    ndev = torch.cuda.device_count()
    devices = [0, 1] if ndev >= 2 else [0, 0]
    for builder, rows, cols in [("kruskal", 1023, 515), ("wilson", 2047, 1025)]:
        maze = lab.build_maze(builder, rows, cols, seed=43)
        after, res, used = lab.distance_map_sharded(maze, devices)
        g_ref, d_ref, mx_ref, settled_ref = O.distance_map(maze)
        assert used == 2
        assert (res.dist == d_ref).all() and (after == g_ref).all() and res.max == mx_ref and res.settled == settled_ref
    # a maze the sharded road does not apply to: one device searches it
    maze = lab.build_maze("grid", 1023, 515, seed=43, mod="cross")
    after, res, used = lab.distance_map_sharded(maze, devices)
    g_ref, d_ref, mx_ref, settled_ref = O.distance_map(maze)
    assert used == 1 and (res.dist == d_ref).all() and (after == g_ref).all()
