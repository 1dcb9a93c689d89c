"""GPU parity: the CUDA path (through the C ABI) against the oracle on identical seeded mazes.

Bit-exact for everything (levels, Square grids, winners, claims, canonical paths)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

BUILDERS = ["arena", "grid", "rdfs", "kruskal", "wilson"]
SIZES = [(7, 7), (9, 13), (51, 111), (63, 65), (129, 127), (255, 383), (511, 511)]


def _distance_case(lab, O, maze):
    with lab.Grid.from_host(maze) as g:
        r = g.distance_map()
        after = g.download()
        stats = g.stats()
    g_ref, d_ref, max_ref, settled_ref = O.distance_map(maze)
    assert (r.dist == d_ref).all(), f"{int((r.dist != d_ref).sum())} levels differ; stats={stats}"
    assert (after == g_ref).all()
    assert r.max == max_ref and r.settled == settled_ref
    return stats


@pytest.mark.parametrize("builder", BUILDERS)
@pytest.mark.parametrize("mod", [None, "cross", "x"])
def test_distance_map_ladder(lab, O, builder, mod):
    for i, (rows, cols) in enumerate(SIZES):
        maze = lab.build_maze(builder, rows, cols, seed=0xB2000000 + i, mod=mod)
        _distance_case(lab, O, maze)


@pytest.mark.parametrize("builder", ["kruskal", "wilson", "grid"])
def test_distance_map_config2_size(lab, O, builder):
    """BASELINE config 2 size: 4095 x 4095."""
    maze = lab.build_maze(builder, 4095, 4095, seed=0xB2000002)
    stats = _distance_case(lab, O, maze)
    if builder != "grid":
        assert stats["resettles"] == 0  # spanning tree: no square is ever settled twice


def test_distance_respects_existing_measure_bits(lab, O):
    """Squares that already carry the measure bit block the search (distance.cc:145-148)."""
    maze = lab.build_maze("arena", 65, 131, seed=1)
    maze[10:40, 70] |= lab.MEASURE_BIT
    maze[30, 20:60] |= lab.MEASURE_BIT
    _distance_case(lab, O, maze)


def test_distance_start_on_wall(lab, O):
    """The start is expanded whatever its bits (distance.cc:135-138)."""
    maze = lab.build_maze("arena", 33, 33, seed=1)
    r, c = lab.distance_start(33, 33)
    maze[r, c] &= ~np.uint16(lab.PATH_BIT)
    _distance_case(lab, O, maze)


def test_repeated_searches_on_one_grid(lab, O):
    maze = lab.build_maze("kruskal", 201, 301, seed=5)
    with lab.Grid.from_host(maze) as g:
        first = g.distance_map()
        g.upload(maze)
        second = g.distance_map()
    assert (first.dist == second.dist).all() and first.max == second.max


def _paths_equal(a, b):
    return len(a) == len(b) and all(x.shape == y.shape and (x == y).all() for x, y in zip(a, b))


@pytest.mark.parametrize("builder", BUILDERS)
@pytest.mark.parametrize("mod", [None, "cross"])
def test_hunt_gather_corners_ladder(lab, O, builder, mod):
    for i, (rows, cols) in enumerate(SIZES[1:]):
        maze = lab.build_maze(builder, rows, cols, seed=0xB2000010 + i, mod=mod)
        # hunt
        placed, s, f = lab.place_hunt(maze, seed=100 + i)
        with lab.Grid.from_host(placed) as g:
            res = g.hunt(s, f)
            after = g.download()
        g_ref, w_ref, p_ref = O.canon_hunt(placed, s, f)
        assert (after == g_ref).all(), ("hunt", rows, cols, int((after != g_ref).sum()))
        assert res.winner == w_ref and _paths_equal(res.paths, [p_ref])
        # gather
        placed, s, fin = lab.place_gather(maze, seed=200 + i)
        with lab.Grid.from_host(placed) as g:
            res = g.gather(s, fin)
            after = g.download()
        g_ref, c_ref, p_ref = O.canon_gather(placed, s, fin)
        assert (after == g_ref).all(), ("gather", rows, cols, int((after != g_ref).sum()))
        assert (res.claimed_by == c_ref).all() and _paths_equal(res.paths, p_ref)
        # corners
        placed, st, f = lab.place_corners(maze, seed=300 + i)
        with lab.Grid.from_host(placed) as g:
            res = g.corners(st, f)
            after = g.download()
        g_ref, w_ref, p_ref = O.canon_corners(placed, st, f)
        assert (after == g_ref).all(), ("corners", rows, cols, int((after != g_ref).sum()))
        assert res.winner == w_ref and _paths_equal(res.paths, [p_ref])


def test_gather_arena_medium(lab, O):
    maze = lab.build_maze("arena", 1023, 1535, seed=1)
    placed, s, fin = lab.place_gather(maze, seed=0xB2000003)
    with lab.Grid.from_host(placed) as g:
        res = g.gather(s, fin, want_paths=False)
        after = g.download()
    g_ref, c_ref, p_ref = O.canon_gather(placed, s, fin)
    assert (after == g_ref).all() and (res.claimed_by == c_ref).all()
    assert res.path_len == [len(p) for p in p_ref]


def test_errors_are_loud(lab):
    with pytest.raises(lab.LabError):
        lab.Grid.empty(2, 2)
    with lab.Grid.empty(9, 9) as g:
        with pytest.raises(lab.LabError):
            g.distance_map(start=(99, 0))


# ---- BASELINE.json's configs at their own sizes, bit-exact against the oracle's CPU restatement ----
def test_config3_gather_arena_16383_bit_exact(lab, O):
    """configs[2]: bfs-gather (4 thread colours) on a 16383 x 16383 arena -- grid, claims and path lengths."""
    maze = lab.cached_maze("arena", 16383, 16383, seed=0xB2000002)
    placed, s, fin = lab.place_gather(maze, seed=0xB2000003)
    with lab.Grid.from_host(placed) as g:
        res = g.gather(s, fin, want_paths=False)
        after = g.download()
    g_ref, c_ref, p_ref = O.canon_gather(placed, s, fin)
    assert (after == g_ref).all(), int((after != g_ref).sum())
    assert (res.claimed_by == c_ref).all() and res.path_len == [len(p) for p in p_ref]


def test_config4_corners_grid_cross_4095_bit_exact(lab, O):
    """configs[3]'s game and builder (bfs-corners, grid -m cross) at 4095 x 4095: grid, winner, the winner's path."""
    maze = lab.cached_maze("grid", 4095, 4095, seed=0xB2000002, mod="cross")
    placed, st, f = lab.place_corners(maze, seed=0xB2000003)
    with lab.Grid.from_host(placed) as g:
        res = g.corners(st, f)
        after = g.download()
    g_ref, w_ref, p_ref = O.canon_corners(placed, st, f)
    assert (after == g_ref).all(), int((after != g_ref).sum())
    assert res.winner == w_ref and _paths_equal(res.paths, [p_ref])


def test_config2_gather_and_hunt_kruskal_4095_bit_exact(lab, O):
    """configs[1]'s maze with the solvers on it (starts on cells and on passages: both take the lattice road)."""
    maze = lab.cached_maze("kruskal", 4095, 4095, seed=0xB2000002)
    for k in range(3):
        placed, s, fin = lab.place_gather(maze, seed=0xB2000003 + k)
        with lab.Grid.from_host(placed) as g:
            res = g.gather(s, fin, want_paths=False)
            after = g.download()
            assert g.stats()["tree_used"] == 3
        g_ref, c_ref, p_ref = O.canon_gather(placed, s, fin)
        assert (after == g_ref).all() and (res.claimed_by == c_ref).all() and res.path_len == [len(p) for p in p_ref]
        placed, s, f = lab.place_hunt(maze, seed=0xB2000013 + k)
        with lab.Grid.from_host(placed) as g:
            res = g.hunt(s, f)
            after = g.download()
        g_ref, w_ref, p_ref = O.canon_hunt(placed, s, f)
        assert (after == g_ref).all() and res.winner == w_ref and _paths_equal(res.paths, [p_ref])
