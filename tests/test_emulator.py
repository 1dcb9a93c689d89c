"""The product's kernel BODIES (csrc/device/engine.cuh, post.cuh) executed on the CPU warp emulator
and compared with the oracle -- the same source nvcc compiles for sm_100a, so the search logic
(bit-sliced batches, seeds, label-correcting min-stores, wake protocol, bounds) is exercised on a
machine without a GPU, under several warp counts and seeded random interleavings."""
import numpy as np
import pytest

from _util import Emu, paths_equal

CASES = [("arena", 63, 65, None), ("arena", 129, 131, None), ("grid", 127, 129, "cross"), ("rdfs", 129, 127, None),
         ("kruskal", 131, 197, None), ("wilson", 129, 129, "cross"), ("kruskal", 9, 13, None)]


@pytest.fixture(scope="module")
def emu():
    return Emu()


@pytest.mark.parametrize("builder,rows,cols,mod", CASES)
def test_emulated_distance_map(lab, O, emu, builder, rows, cols, mod):
    maze = lab.build_maze(builder, rows, cols, seed=0xE0 + rows, mod=mod)
    g_ref, d_ref, mx_ref, settled_ref = O.distance_map(maze)
    for nwarps, sched in [(1, 0), (4, 0), (7, 3), (16, 9)]:
        g, d, mx, settled, _ = emu.distance(maze, nwarps=nwarps, sched=sched)
        assert (d == d_ref).all() and (g == g_ref).all() and mx == mx_ref and settled == settled_ref, (nwarps, sched)


@pytest.mark.parametrize("builder,rows,cols,mod", CASES[:6])
def test_emulated_games(lab, O, emu, builder, rows, cols, mod):
    maze = lab.build_maze(builder, rows, cols, seed=0xE1 + cols, mod=mod)
    for nwarps, sched in [(4, 0), (9, 7)]:
        placed, s, f = lab.place_hunt(maze, seed=50)
        g_ref, w_ref, p_ref = O.canon_hunt(placed, s, f)
        g, w, _, paths, _ = emu.solver(1, placed, [s], [f], nwarps, sched)
        assert (g == g_ref).all() and w == w_ref and paths_equal([paths[0]], [p_ref])
        placed, s, fin = lab.place_gather(maze, seed=60)
        g_ref, c_ref, p_ref = O.canon_gather(placed, s, fin)
        g, _, c, paths, _ = emu.solver(2, placed, [s], fin, nwarps, sched)
        assert (g == g_ref).all() and (c == c_ref).all() and paths_equal(paths, p_ref)
        placed, st, f = lab.place_corners(maze, seed=70)
        g_ref, w_ref, p_ref = O.canon_corners(placed, st, f)
        g, w, _, paths, _ = emu.solver(3, placed, st, [f], nwarps, sched)
        assert (g == g_ref).all() and w == w_ref and paths_equal([paths[0]], [p_ref])
