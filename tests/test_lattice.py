"""The lattice road (csrc/device/lattice.cuh): the reference builders' perfect mazes as spanning trees of the lattice
of odd squares -- bit-board flood fills per tile, the Euler tour from the clockwise order of a component's crossings,
hierarchical in-place pointer jumping, a prefix sum over the tour.  It must (a) be the road the reference builders'
mazes take, (b) give exactly the FIFO search's levels (painters/distance.cc:129-153), (c) step aside, having written
nothing, for everything else: trees of squares that are not on the lattice (-> the general pipeline, tree.cuh) and
lattices that only LOOK like trees (-> the tile wave, engine.cuh).  CPU half: the kernel bodies on the warp emulator;
GPU half: through the C ABI."""
import os

import numpy as np
import pytest

from _util import Emu, paths_equal
from test_tree_pipeline import tree_looking_non_tree

LATTICE, GENERAL, WAVE = 3, 1, 0


@pytest.fixture(scope="module")
def emu():
    return Emu()


LADDER = [("kruskal", 7, 7), ("kruskal", 9, 63), ("kruskal", 65, 67), ("wilson", 127, 193), ("rdfs", 129, 191), ("wilson", 63, 65),
          ("kruskal", 255, 257), ("rdfs", 65, 513)]


@pytest.mark.parametrize("builder,rows,cols", LADDER)
def test_emulated_lattice_road_is_used_and_exact(lab, O, emu, builder, rows, cols):
    maze = lab.build_maze(builder, rows, cols, seed=0x1A + cols)
    g_ref, d_ref, mx_ref, settled_ref = O.distance_map(maze)
    for nwarps, sched in [(1, 0), (5, 0), (7, 13)]:
        g, d, mx, settled, _ = emu.distance(maze, nwarps=nwarps, sched=sched)
        assert emu.lib.emu_last_tree_used() == LATTICE and emu.lib.emu_last_lattice_error() == 0
        assert (d == d_ref).all() and (g == g_ref).all() and mx == mx_ref and settled == settled_ref, (nwarps, sched)


def test_emulated_several_super_tiles(lab, O, emu):
    """More than LST = 8 tiles a side: the local ranking leaves border arcs to the global one."""
    maze = lab.build_maze("kruskal", 641, 1089, seed=3)  # 11 x 18 tiles: 2 x 3 super-tiles
    g_ref, d_ref, mx_ref, settled_ref = O.distance_map(maze)
    g, d, mx, settled, _ = emu.distance(maze, nwarps=6, sched=5)
    assert emu.lib.emu_last_tree_used() == LATTICE
    assert (d == d_ref).all() and (g == g_ref).all() and mx == mx_ref and settled == settled_ref


def test_emulated_three_roads_agree(lab, O, emu):
    maze = lab.build_maze("wilson", 193, 129, seed=77)
    out = {}
    try:
        for road, (tree, lattice) in {LATTICE: (1, 1), GENERAL: (1, 0), WAVE: (0, 0)}.items():
            emu.lib.emu_set_tree(tree)
            emu.lib.emu_set_lattice(lattice)
            out[road] = emu.distance(maze, nwarps=6, sched=3)
            assert emu.lib.emu_last_tree_used() == road
    finally:
        emu.lib.emu_set_tree(1)
        emu.lib.emu_set_lattice(1)
    for road in (GENERAL, WAVE):
        assert (out[road][1] == out[LATTICE][1]).all() and (out[road][0] == out[LATTICE][0]).all()
        assert out[road][2:4] == out[LATTICE][2:4]


@pytest.mark.parametrize("start", [(1, 1), (63, 63), (65, 63), (63, 65), (127, 191), (65, 1), (1, 191), (33, 77)])
def test_emulated_any_odd_root(lab, O, emu, start):
    """Roots in tile corners, next to tile borders, in the maze's corners."""
    maze = lab.build_maze("kruskal", 129, 193, seed=5)
    d_ref = O.bfs_levels(maze, *start)
    g, d, mx, settled, _ = emu.distance(maze, start=start, nwarps=4, sched=2)
    assert emu.lib.emu_last_tree_used() == LATTICE
    reached = d_ref != 0xFFFFFFFF
    assert (d == d_ref).all() and mx == int(d_ref[reached].max()) and settled == int(reached.sum())


def passage_squares(lab, maze):
    """(odd, even) and (even, odd) path squares: next to tile borders, ON tile borders (column / row 64, 128), in corners."""
    rows, cols = maze.shape
    want = [(1, 2), (2, 1), (63, 64), (64, 63), (65, 64), (64, 65), (63, 62), (62, 63), (127, 128), (128, 127), (rows - 2, cols - 3), (rows - 3, cols - 2),
            (33, 100), (100, 33)]
    out = []
    for r0, c0 in want:
        # the nearest passable square with the same parities
        for dr in range(0, rows, 2):
            hit = [(r, c) for r in (r0 - dr, r0 + dr) for dc in range(0, 40, 2) for c in (c0 - dc, c0 + dc)
                   if 0 < r < rows - 1 and 0 < c < cols - 1 and maze[r, c] & lab.PATH_BIT]
            if hit:
                out.append(hit[0])
                break
    return sorted(set(out))


def test_emulated_roots_on_passages(lab, O, emu):
    """The solvers start anywhere (solve_utilities.cc:275-285): a start on a passage square splits the lattice into two
    trees whose tours are chained."""
    maze = lab.build_maze("kruskal", 131, 195, seed=6)
    starts = passage_squares(lab, maze)
    assert len(starts) >= 10 and any(c % 64 == 0 for r, c in starts) and any(r % 64 == 0 for r, c in starts)
    for k, (r, c) in enumerate(starts):
        assert (r + c) % 2 == 1
        d_ref = O.bfs_levels(maze, r, c)
        g, d, mx, settled, _ = emu.distance(maze, start=(r, c), nwarps=4, sched=k)
        assert emu.lib.emu_last_tree_used() == LATTICE, (r, c)
        reached = d_ref != 0xFFFFFFFF
        assert (d == d_ref).all() and mx == int(d_ref[reached].max()) and settled == int(reached.sum()), (r, c)
        assert (((g & lab.MEASURE_BIT) != 0) == reached).all()


def test_emulated_root_on_a_passage_inside_one_tile(lab, O, emu):
    maze = lab.build_maze("wilson", 31, 41, seed=2)
    for (r, c) in passage_squares(lab, maze)[:6]:
        d_ref = O.bfs_levels(maze, r, c)
        g, d, mx, settled, _ = emu.distance(maze, start=(r, c), nwarps=2, sched=1)
        assert emu.lib.emu_last_tree_used() == LATTICE
        assert (d == d_ref).all()


def off_lattice_tree(lab, rows, cols, seed):
    """A perfect maze with one pillar (even, even) opened next to a passage: still a tree of squares (a leaf more),
    no longer a lattice."""
    maze = lab.build_maze("kruskal", rows, cols, seed=seed)
    path = (maze & lab.PATH_BIT) != 0
    # pillars sit at (even, even)
    for r in range(2, rows - 2, 2):
        for c in range(2, cols - 2, 2):
            if int(path[r - 1, c]) + int(path[r + 1, c]) + int(path[r, c - 1]) + int(path[r, c + 1]) == 1:
                maze[r, c] |= lab.PATH_BIT
                return maze
    raise AssertionError("no pillar with exactly one passable neighbour")


@pytest.mark.parametrize("seed", [1, 2])
def test_emulated_off_lattice_tree_takes_the_general_pipeline(lab, O, emu, seed):
    maze = off_lattice_tree(lab, 129, 131, seed)
    g_ref, d_ref, mx_ref, settled_ref = O.distance_map(maze)
    g, d, mx, settled, _ = emu.distance(maze, nwarps=4, sched=seed)
    assert emu.lib.emu_last_lattice_error() == 1 and emu.lib.emu_last_tree_used() == GENERAL
    assert (d == d_ref).all() and (g == g_ref).all() and mx == mx_ref and settled == settled_ref


@pytest.mark.parametrize("seed", [1, 2, 3, 4, 5])
def test_emulated_lattice_that_only_looks_like_a_tree_falls_back(lab, O, emu, seed):
    """One wall opened (a cycle), one passage closed (a part cut off): a lattice with a tree's counts.  The tour breaks
    into several cycles or the floods miss squares; both roads for trees give up and the wave searches."""
    maze = tree_looking_non_tree(lab, 129, 131, seed)
    g_ref, d_ref, mx_ref, settled_ref = O.distance_map(maze)
    g, d, mx, settled, _ = emu.distance(maze, nwarps=4, sched=seed)
    assert emu.lib.emu_last_lattice_error() in (3, 4, 5) and emu.lib.emu_last_tree_used() == WAVE
    assert (d == d_ref).all() and (g == g_ref).all() and mx == mx_ref and settled == settled_ref


def test_emulated_games_on_the_lattice(lab, O, emu):
    maze = lab.build_maze("kruskal", 131, 197, seed=9)
    seen = set()
    for seed in range(50, 58):
        placed, s, f = lab.place_hunt(maze, seed=seed)
        g_ref, w_ref, p_ref = O.canon_hunt(placed, s, f)
        g, w, _, paths, _ = emu.solver(1, placed, [s], [f], 4, 1)
        assert emu.lib.emu_last_tree_used() == LATTICE and emu.lib.emu_last_paint_fused() == 1  # (paint rules applied by the levels kernel)
        assert emu.lib.emu_last_path_segments() >= 1 or len(p_ref) <= 1  # (the winner's path came from the tour's ranks, in segments)
        seen.add((s[0] % 2, s[1] % 2))
        assert (g == g_ref).all() and w == w_ref and paths_equal([paths[0]], [p_ref])
    assert len(seen) >= 2  # starts on cells and on passages
    for seed in range(60, 64):
        placed, s, fin = lab.place_gather(maze, seed=seed)
        g_ref, c_ref, p_ref = O.canon_gather(placed, s, fin)
        g, _, c, paths, _ = emu.solver(2, placed, [s], fin, 4, 1)
        assert emu.lib.emu_last_paint_fused() == 1 and emu.lib.emu_last_path_segments() >= 4  # (all four paths from the tour's ranks, in segments)
        assert (g == g_ref).all() and (c == c_ref).all() and paths_equal(paths, p_ref)


def test_emulated_gather_without_its_field(lab, O, emu):
    """LAB_OPT_SOLVER_LEVELS off, no paths wanted: bfs-gather on the lattice road applies its paint rules WITHOUT writing the
    level field (SURVEY 8(d)'s 4 bytes per square); path.front() of every colour comes from probing the finishes'
    neighbours.  The grid is the canonical model's all the same."""
    maze = lab.build_maze("kruskal", 131, 197, seed=9)
    emu.lib.emu_set_option_solver_levels(0)
    try:
        for seed in range(60, 70):
            placed, s, fin = lab.place_gather(maze, seed=seed)
            g_ref, c_ref, p_ref = O.canon_gather(placed, s, fin)
            g, _, c, _, painted = emu.solver(2, placed, [s], fin, 4, 1, want_paths=False)
            assert emu.lib.emu_last_tree_used() == LATTICE and emu.lib.emu_last_paint_fused() == 2  # (2: painted, field skipped)
            assert (g == g_ref).all() and (c == c_ref).all() and painted > 0
    finally:
        emu.lib.emu_set_option_solver_levels(1)


# ------------------------------------------------------------------------------------------- GPU
@pytest.mark.gpu
def test_gpu_gather_without_its_field(lab, O):
    """lab_grid_set_option(LAB_OPT_SOLVER_LEVELS, 0): the same grid and claims, no field (lab_levels_download says so)."""
    maze = lab.build_maze("wilson", 1023, 1535, seed=4)
    for seed in range(4):
        placed, s, fin = lab.place_gather(maze, seed=0xB2000203 + seed)
        g_ref, c_ref, p_ref = O.canon_gather(placed, s, fin)
        with lab.Grid.from_host(placed) as g:
            g.set_option(lab.OPT_SOLVER_LEVELS, 0)
            res = g.gather(s, fin, want_paths=False)
            assert g.stats()["tree_used"] == LATTICE
            assert (g.download() == g_ref).all() and (res.claimed_by == c_ref).all() and list(res.path_len) == [len(p) for p in p_ref]
            with pytest.raises(lab.LabError):
                g.levels(0)
            # with paths the field is needed and written whatever the option says
            g.upload(placed)
            res = g.gather(s, fin, want_paths=True)
            assert (g.download() == g_ref).all() and paths_equal(res.paths, p_ref)
            g.levels(0)

@pytest.mark.gpu
@pytest.mark.parametrize("builder,n", [("kruskal", 1023), ("wilson", 2047), ("rdfs", 1535), ("kruskal", 4095), ("wilson", 4095)])
def test_gpu_lattice_road_is_used_and_exact(lab, O, builder, n):
    maze = lab.build_maze(builder, n, n, seed=0xB2000002)
    with lab.Grid.from_host(maze) as g:
        r = g.distance_map()
        after = g.download()
        stats = g.stats()
    g_ref, d_ref, max_ref, settled_ref = O.distance_map(maze)
    assert stats["tree_used"] == LATTICE and stats["tree_pairs"] + 1 == stats["tree_squares"] == settled_ref
    assert (r.dist == d_ref).all() and (after == g_ref).all() and r.max == max_ref and r.settled == settled_ref


@pytest.mark.gpu
def test_gpu_lattice_rectangular_and_ragged_sizes(lab, O):
    for rows, cols in [(65, 4097), (3001, 129), (1023, 1025), (769, 63), (7, 7), (513, 515)]:
        maze = lab.build_maze("wilson", rows, cols, seed=rows)
        with lab.Grid.from_host(maze) as g:
            r = g.distance_map()
            stats = g.stats()
        g_ref, d_ref, max_ref, settled_ref = O.distance_map(maze)
        assert stats["tree_used"] == LATTICE
        assert (r.dist == d_ref).all() and r.max == max_ref and r.settled == settled_ref, (rows, cols)


@pytest.mark.gpu
def test_gpu_repeated_searches_on_one_grid(lab, O):
    """The workspace is reused: a second and third search (other roots) on the same grid."""
    maze = lab.build_maze("kruskal", 1023, 769, seed=12)
    with lab.Grid.from_host(maze) as g:
        for start in [(511, 385), (1, 1), (1021, 767), (3, 765)]:
            g.upload(maze)
            r = g.distance_map(start)
            assert g.stats()["tree_used"] == LATTICE
            assert (r.dist == O.bfs_levels(maze, *start)).all(), start


@pytest.mark.gpu
@pytest.mark.parametrize("seed", [1, 2, 3, 4])
def test_gpu_fallbacks(lab, O, seed):
    for maze, road in [(tree_looking_non_tree(lab, 511, 767, seed), WAVE), (off_lattice_tree(lab, 511, 767, seed), GENERAL)]:
        with lab.Grid.from_host(maze) as g:
            r = g.distance_map()
            after = g.download()
            stats = g.stats()
        g_ref, d_ref, max_ref, settled_ref = O.distance_map(maze)
        assert stats["tree_used"] == road
        assert (r.dist == d_ref).all() and (after == g_ref).all() and r.max == max_ref and r.settled == settled_ref


@pytest.mark.gpu
def test_gpu_games_on_the_lattice(lab, O):
    maze = lab.build_maze("kruskal", 1023, 1535, seed=3)
    roads = set()
    for seed in range(8):
        placed, s, fin = lab.place_gather(maze, seed=0xB2000003 + seed)
        with lab.Grid.from_host(placed) as g:
            res = g.gather(s, fin)
            after = g.download()
            roads.add(g.stats()["tree_used"])
        g_ref, c_ref, p_ref = O.canon_gather(placed, s, fin)
        assert (after == g_ref).all() and (res.claimed_by == c_ref).all() and paths_equal(res.paths, p_ref)
        placed, s, f = lab.place_hunt(maze, seed=0xB2000103 + seed)
        with lab.Grid.from_host(placed) as g:
            res = g.hunt(s, f)
            after = g.download()
            roads.add(g.stats()["tree_used"])
        g_ref, w_ref, p_ref = O.canon_hunt(placed, s, f)
        assert (after == g_ref).all() and res.winner == w_ref and paths_equal(res.paths, [p_ref])
    assert roads == {LATTICE}


@pytest.mark.gpu
def test_gpu_headline_size_32767(lab, O):
    """BASELINE north_star size, one GPU, bit-exact against the oracle's CPU restatement (cached maze file)."""
    n = 32767
    maze = lab.cached_maze("kruskal", n, n, seed=0xB2000002)
    with lab.Grid.from_host(maze) as g:
        r = g.distance_map()
        stats = g.stats()
        t = g.timing()
    g_ref, d_ref, max_ref, settled_ref = O.distance_map(maze)
    assert stats["tree_used"] == LATTICE and r.settled == settled_ref == 2 * 16383 * 16383 - 1 and r.max == max_ref
    assert (r.dist == d_ref).all()
    print(f"32767^2 kruskal: max level {r.max}, search {t['search_ms']:.2f} ms, total {t['total_ms']:.2f} ms")
