"""The spanning-tree pipeline (csrc/device/tree.cuh): Euler-tour walk, pointer jumping, independent tile
floods.  It must (a) be the path that runs on the reference builders' perfect mazes, (b) give exactly
the levels of the FIFO search (painters/distance.cc:129-153) and (c) step aside -- leaving the tile wave
to do the search -- on anything that is not one tree, including mazes whose square / adjacency counts
LOOK like a tree's.  CPU half: the same kernel bodies on the warp emulator; GPU half: through the C ABI."""
import os

import numpy as np
import pytest

from _util import Emu, paths_equal

PERFECT = [("kruskal", 65, 67), ("kruskal", 127, 193), ("wilson", 131, 129), ("rdfs", 129, 191), ("kruskal", 255, 257)]


@pytest.fixture(scope="module")
def emu():
    return Emu()


@pytest.fixture(autouse=True)
def general_pipeline_only(emu):
    """These tests are about the GENERAL pipeline (any tree of squares): the lattice road (device/lattice.cuh,
    tests/test_lattice.py), which would take the reference builders' mazes first, is switched off."""
    emu.lib.emu_set_lattice(0)
    old = os.environ.get("LAB_LATTICE")
    os.environ["LAB_LATTICE"] = "0"
    yield
    emu.lib.emu_set_lattice(1)
    if old is None:
        del os.environ["LAB_LATTICE"]
    else:
        os.environ["LAB_LATTICE"] = old


def tree_looking_non_tree(lab, rows, cols, seed):
    """A perfect maze with one wall opened (a cycle) and one passage closed elsewhere (a part cut off):
    still squares - 1 adjacent pairs, but neither connected nor acyclic."""
    maze = lab.build_maze("kruskal", rows, cols, seed=seed)
    path = (maze & lab.PATH_BIT) != 0
    rng = np.random.default_rng(seed)
    # open a wall between two passable squares that are two apart in a row (creates a cycle)
    cand = [(r, c) for r in range(1, rows - 1, 2) for c in range(2, cols - 2, 2) if not path[r, c] and path[r, c - 1] and path[r, c + 1]]
    r, c = cand[rng.integers(len(cand))]
    maze[r, c] |= lab.PATH_BIT
    # close a passage square far from the start that has exactly two passable neighbours (cuts a part off)
    path = (maze & lab.PATH_BIT) != 0
    cand = [(r2, c2) for r2 in range(1, rows - 1) for c2 in range(1, cols - 1)
            if path[r2, c2] and (r2 % 2 != c2 % 2) and abs(r2 - rows // 2) + abs(c2 - cols // 2) > 20 and (r2, c2) != (r, c)
            and abs(r2 - r) + abs(c2 - c) > 4]
    r2, c2 = cand[rng.integers(len(cand))]
    maze[r2, c2] &= ~np.uint16(lab.PATH_BIT)
    return maze


@pytest.mark.parametrize("builder,rows,cols", PERFECT)
def test_emulated_tree_pipeline_is_used_and_exact(lab, O, emu, builder, rows, cols):
    maze = lab.build_maze(builder, rows, cols, seed=0x7E + cols)
    g_ref, d_ref, mx_ref, settled_ref = O.distance_map(maze)
    for nwarps, sched in [(1, 0), (5, 0), (8, 11)]:
        g, d, mx, settled, _ = emu.distance(maze, nwarps=nwarps, sched=sched)
        assert emu.lib.emu_last_tree_used() == 1
        assert (d == d_ref).all() and (g == g_ref).all() and mx == mx_ref and settled == settled_ref, (nwarps, sched)


def test_emulated_tree_and_wave_agree(lab, O, emu):
    maze = lab.build_maze("wilson", 193, 129, seed=77)
    try:
        emu.lib.emu_set_tree(0)
        g0, d0, mx0, st0, _ = emu.distance(maze, nwarps=6, sched=3)
        assert emu.lib.emu_last_tree_used() == 0
    finally:
        emu.lib.emu_set_tree(1)
    g1, d1, mx1, st1, _ = emu.distance(maze, nwarps=6, sched=3)
    assert emu.lib.emu_last_tree_used() == 1
    assert (d0 == d1).all() and (g0 == g1).all() and mx0 == mx1 and st0 == st1


@pytest.mark.parametrize("start", [(1, 1), (63, 64), (64, 63), (127, 191), (65, 1)])
def test_emulated_tree_any_root(lab, O, emu, start):
    """Roots on tile borders and corners (the tour starts at the root: tree.cuh walk_tile)."""
    maze = lab.build_maze("kruskal", 129, 193, seed=5)
    r, c = start
    if not maze[r, c] & lab.PATH_BIT:  # take the nearest passable square
        r, c = next((rr, cc) for rr in (r, r - 1, r + 1) for cc in (c, c - 1, c + 1) if maze[rr, cc] & lab.PATH_BIT)
    d_ref = O.bfs_levels(maze, r, c)  # plain BFS over path squares from (r, c)
    g, d, mx, settled, _ = emu.distance(maze, start=(r, c), nwarps=4, sched=2)
    assert emu.lib.emu_last_tree_used() == 1
    reached = d_ref != 0xFFFFFFFF
    assert (d == d_ref).all() and mx == int(d_ref[reached].max()) and settled == int(reached.sum())


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_emulated_non_tree_with_tree_counts_falls_back(lab, O, emu, seed):
    maze = tree_looking_non_tree(lab, 129, 131, seed)
    g_ref, d_ref, mx_ref, settled_ref = O.distance_map(maze)
    g, d, mx, settled, _ = emu.distance(maze, nwarps=4, sched=seed)
    assert emu.lib.emu_last_tree_used() == 0  # counts say "tree", the checks say no: the wave did it
    assert (d == d_ref).all() and (g == g_ref).all() and mx == mx_ref and settled == settled_ref


def test_emulated_games_on_trees_use_the_pipeline(lab, O, emu):
    maze = lab.build_maze("kruskal", 131, 197, seed=9)
    placed, s, f = lab.place_hunt(maze, seed=50)
    g_ref, w_ref, p_ref = O.canon_hunt(placed, s, f)
    g, w, _, paths, _ = emu.solver(1, placed, [s], [f], 4, 1)
    assert emu.lib.emu_last_tree_used() == 1
    assert (g == g_ref).all() and w == w_ref and paths_equal([paths[0]], [p_ref])
    placed, s, fin = lab.place_gather(maze, seed=60)
    g_ref, c_ref, p_ref = O.canon_gather(placed, s, fin)
    g, _, c, paths, _ = emu.solver(2, placed, [s], fin, 4, 1)
    assert emu.lib.emu_last_tree_used() == 1
    assert (g == g_ref).all() and (c == c_ref).all() and paths_equal(paths, p_ref)


def test_emulated_long_winner_path_is_walked_in_batches(lab, O, emu):
    """rdfs: one corridor thousands of squares long.  The path walk (post.cuh walk_body) caches a tile of levels at a
    time and flushes the collected squares 1024 at a time: several tiles, several flushes, same path and repaint."""
    maze = lab.build_maze("rdfs", 131, 197, seed=4)
    placed, s, f = lab.place_hunt(maze, seed=52)
    g_ref, w_ref, p_ref = O.canon_hunt(placed, s, f)
    assert len(p_ref) > 1100, len(p_ref)
    g, w, _, paths, _ = emu.solver(1, placed, [s], [f], 4, 1)
    assert (g == g_ref).all() and w == w_ref and paths_equal([paths[0]], [p_ref])


# ------------------------------------------------------------------------------------------- GPU
@pytest.mark.gpu
@pytest.mark.parametrize("builder,n", [("kruskal", 1023), ("wilson", 2047), ("rdfs", 1535), ("kruskal", 4095)])
def test_gpu_tree_pipeline_is_used_and_exact(lab, O, builder, n):
    maze = lab.build_maze(builder, n, n, seed=0xB2000002)
    with lab.Grid.from_host(maze) as g:
        r = g.distance_map()
        after = g.download()
        stats = g.stats()
    g_ref, d_ref, max_ref, settled_ref = O.distance_map(maze)
    assert stats["tree_used"] == 1 and stats["tree_pairs"] + 1 == stats["tree_squares"] == settled_ref
    assert (r.dist == d_ref).all() and (after == g_ref).all() and r.max == max_ref and r.settled == settled_ref


@pytest.mark.gpu
@pytest.mark.parametrize("seed", [1, 2, 3, 4])
def test_gpu_non_tree_with_tree_counts_falls_back(lab, O, seed):
    maze = tree_looking_non_tree(lab, 511, 767, seed)
    with lab.Grid.from_host(maze) as g:
        r = g.distance_map()
        after = g.download()
        stats = g.stats()
    g_ref, d_ref, max_ref, settled_ref = O.distance_map(maze)
    assert stats["tree_pairs"] + 1 == stats["tree_squares"] and stats["tree_used"] == 0
    assert (r.dist == d_ref).all() and (after == g_ref).all() and r.max == max_ref and r.settled == settled_ref


@pytest.mark.gpu
def test_gpu_tree_rectangular_and_ragged_sizes(lab, O):
    for rows, cols in [(65, 4097), (3001, 129), (1023, 1025), (769, 63)]:
        maze = lab.build_maze("wilson", rows, cols, seed=rows)
        with lab.Grid.from_host(maze) as g:
            r = g.distance_map()
            stats = g.stats()
        g_ref, d_ref, max_ref, settled_ref = O.distance_map(maze)
        assert stats["tree_used"] == 1
        assert (r.dist == d_ref).all() and r.max == max_ref and r.settled == settled_ref, (rows, cols)


@pytest.mark.gpu
def test_gpu_gather_on_a_tree_uses_the_pipeline(lab, O):
    maze = lab.build_maze("kruskal", 1023, 1535, seed=3)
    placed, s, fin = lab.place_gather(maze, seed=0xB2000003)
    with lab.Grid.from_host(placed) as g:
        res = g.gather(s, fin)
        after = g.download()
        stats = g.stats()
    g_ref, c_ref, p_ref = O.canon_gather(placed, s, fin)
    assert stats["tree_used"] == 1
    assert (after == g_ref).all() and (res.claimed_by == c_ref).all() and paths_equal(res.paths, p_ref)


@pytest.mark.gpu
@pytest.mark.skipif(not os.environ.get("LAB_BIG_TESTS"), reason="north_star headline size: ~3 min of host work; set LAB_BIG_TESTS=1")
@pytest.mark.parametrize("builder", ["kruskal"])
def test_gpu_tree_headline_size_32767(lab, O, builder):
    """BASELINE north_star size, one GPU: bit-exact against the oracle's CPU restatement."""
    n = 32767
    maze = lab.build_maze(builder, n, n, seed=0xB2000002)
    with lab.Grid.from_host(maze) as g:
        r = g.distance_map()
        stats = g.stats()
        t = g.timing()
    g_ref, d_ref, max_ref, settled_ref = O.distance_map(maze)
    assert stats["tree_used"] == 1 and r.settled == settled_ref == 2 * 16383 * 16383 - 1 and r.max == max_ref
    assert (r.dist == d_ref).all()
    print(f"32767^2 {builder}: max level {r.max}, search {t['search_ms']:.2f} ms, total {t['total_ms']:.2f} ms")
