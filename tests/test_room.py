"""The open room (csrc/device/room.cuh): a maze whose passable squares are exactly one full rectangle -- the
reference's arena builder (builders/arena.cc:17-26) -- gets its levels from a closed form (Manhattan distance)
in one streaming pass instead of a search.  It must (a) be the path that runs on arenas, (b) give exactly the
levels / Square bits / paths of the FIFO search (painters/distance.cc:129-153, solvers/bfs_threads.cc), and
(c) step aside for anything else: one wall square inside an arena is enough.
CPU half: the same kernel bodies on the warp emulator; GPU half: through the C ABI."""
import numpy as np
import pytest

from _util import Emu, paths_equal

ROOM = 2  # lab_stats.tree_used / emu_last_tree_used: 0 wave, 1 tree pipeline, 2 open room


@pytest.fixture(scope="module")
def emu():
    return Emu()


def boxed_room(lab, rows, cols, r0, r1, c0, c1):
    """Walls everywhere except the full rectangle [r0, r1] x [c0, c1]."""
    maze = np.zeros((rows, cols), np.uint16)
    maze[r0 : r1 + 1, c0 : c1 + 1] = lab.PATH_BIT
    return maze


@pytest.mark.parametrize("rows,cols", [(7, 7), (9, 71), (65, 67), (129, 63), (131, 197)])
def test_emulated_arena_distance_takes_the_room(lab, O, emu, rows, cols):
    maze = lab.build_maze("arena", rows, cols, seed=1)
    g_ref, d_ref, mx_ref, settled_ref = O.distance_map(maze)
    for nwarps, sched in [(1, 0), (5, 3)]:
        g, d, mx, settled, _ = emu.distance(maze, nwarps=nwarps, sched=sched)
        assert emu.lib.emu_last_tree_used() == ROOM
        assert (d == d_ref).all() and (g == g_ref).all() and mx == mx_ref and settled == settled_ref


@pytest.mark.parametrize("start", [(1, 1), (1, 195), (129, 1), (64, 64), (63, 128), (100, 7)])
def test_emulated_room_any_start(lab, O, emu, start):
    maze = lab.build_maze("arena", 131, 197, seed=1)
    d_ref = O.bfs_levels(maze, *start)
    g, d, mx, settled, _ = emu.distance(maze, start=start, nwarps=4, sched=1)
    assert emu.lib.emu_last_tree_used() == ROOM
    reached = d_ref != 0xFFFFFFFF
    assert (d == d_ref).all() and mx == int(d_ref[reached].max()) and settled == int(reached.sum())


@pytest.mark.parametrize("builder", ["arena", "kruskal"])
@pytest.mark.parametrize("misalign", [1, 4, 7])
def test_emulated_square_array_off_the_16_byte_grid(lab, O, emu, builder, misalign):
    """pack_body and room_stream read 16 bytes at a time; a caller's array may start anywhere."""
    maze = lab.build_maze(builder, 67, 133, seed=3)
    g_ref, d_ref, mx_ref, settled_ref = O.distance_map(maze)
    g, d, mx, settled, _ = emu.distance(maze, nwarps=3, sched=1, misalign=misalign)
    assert (d == d_ref).all() and (g == g_ref).all() and mx == mx_ref and settled == settled_ref


def test_emulated_room_smaller_than_the_grid(lab, O, emu):
    """The rectangle need not be the whole interior; squares outside it stay untouched and get no level."""
    maze = boxed_room(lab, 140, 150, 3, 70, 66, 140)
    maze[0, 0] = 0x000F  # wall shapes outside the room are nobody's business
    start = (40, 100)
    d_ref = O.bfs_levels(maze, *start)
    g, d, mx, settled, _ = emu.distance(maze, start=start, nwarps=3, sched=2)
    assert emu.lib.emu_last_tree_used() == ROOM
    assert (d == d_ref).all() and settled == 68 * 75 and mx == int(d_ref[d_ref != 0xFFFFFFFF].max())
    assert ((g & lab.MEASURE_BIT) != 0).sum() == 68 * 75 and g[0, 0] == 0x000F


@pytest.mark.parametrize("hole", [(64, 64), (1, 1), (129, 195), (70, 3)])
def test_emulated_arena_with_one_wall_is_not_a_room(lab, O, emu, hole):
    maze = lab.build_maze("arena", 131, 197, seed=1)
    maze[hole] &= ~np.uint16(lab.PATH_BIT)
    g_ref, d_ref, mx_ref, settled_ref = O.distance_map(maze)
    g, d, mx, settled, _ = emu.distance(maze, nwarps=4, sched=1)
    assert emu.lib.emu_last_tree_used() == 0  # the tile wave did it
    assert (d == d_ref).all() and (g == g_ref).all() and mx == mx_ref and settled == settled_ref


def test_emulated_two_rooms_are_not_a_room(lab, O, emu):
    """Same square count as a box can only happen for the box itself; two rooms have a larger bounding box."""
    maze = boxed_room(lab, 80, 140, 2, 30, 2, 60)
    maze[40:70, 70:130] = lab.PATH_BIT
    start = (10, 10)
    d_ref = O.bfs_levels(maze, *start)
    g, d, mx, settled, _ = emu.distance(maze, start=start, nwarps=4, sched=0)
    assert emu.lib.emu_last_tree_used() == 0
    assert (d == d_ref).all() and settled == 29 * 59


def test_emulated_measured_squares_break_the_room(lab, O, emu):
    """distance.cc:145-148 skips squares that already carry the measure bit: a second map over a measured arena
    only has its start square to settle (a 1 x 1 room), like the reference."""
    maze = lab.build_maze("arena", 67, 69, seed=1)
    g1, _, _, _, _ = emu.distance(maze)
    g_ref, d_ref, mx_ref, settled_ref = O.distance_map(g1)
    g2, d2, mx2, settled2, _ = emu.distance(g1)
    assert (d2 == d_ref).all() and (g2 == g_ref).all() and mx2 == mx_ref == 0 and settled2 == settled_ref == 1


def test_emulated_games_in_a_room(lab, O, emu):
    maze = lab.build_maze("arena", 67, 131, seed=1)
    for seed in (50, 51):
        placed, s, f = lab.place_hunt(maze, seed=seed)
        g_ref, w_ref, p_ref = O.canon_hunt(placed, s, f)
        g, w, _, paths, _ = emu.solver(1, placed, [s], [f], 4, 1)
        assert emu.lib.emu_last_tree_used() == ROOM
        assert (g == g_ref).all() and w == w_ref and paths_equal([paths[0]], [p_ref])
        placed, s, fin = lab.place_gather(maze, seed=seed + 10)
        g_ref, c_ref, p_ref = O.canon_gather(placed, s, fin)
        g, _, c, paths, _ = emu.solver(2, placed, [s], fin, 4, 1)
        assert emu.lib.emu_last_tree_used() == ROOM
        assert (g == g_ref).all() and (c == c_ref).all() and paths_equal(paths, p_ref)


@pytest.mark.parametrize("finish", [(5, 5), (5, 28), (5, 50), (20, 5), (20, 50), (35, 5), (35, 28), (35, 50), (19, 28), (20, 29), (21, 28), (20, 27),
                                    (1, 1), (39, 55)])
def test_emulated_room_paths_are_the_canonical_two_segments(lab, O, emu, finish):
    """In a room the path is written from a closed form (post.cuh walk_body): it must be the oracle's canonical walk
    (first of N, E, S, W one level lower) for every position of the finish relative to the start."""
    maze = lab.build_maze("arena", 41, 57, seed=1)
    start = (20, 28)
    placed = maze.copy()
    placed[start] |= lab.START_BIT
    placed[finish] |= lab.FINISH_BIT
    g_ref, w_ref, p_ref = O.canon_hunt(placed, start, finish)
    g, w, _, paths, _ = emu.solver(1, placed, [start], [finish], 2, 1)
    assert emu.lib.emu_last_tree_used() == ROOM
    assert (g == g_ref).all() and w == w_ref and paths_equal([paths[0]], [p_ref])


# ------------------------------------------------------------------------------------------- GPU
@pytest.mark.gpu
@pytest.mark.parametrize("rows,cols", [(7, 7), (255, 383), (1023, 1025), (2047, 4095)])
def test_gpu_arena_distance_takes_the_room(lab, O, rows, cols):
    maze = lab.build_maze("arena", rows, cols, seed=1)
    with lab.Grid.from_host(maze) as g:
        r = g.distance_map()
        after = g.download()
        stats = g.stats()
    g_ref, d_ref, max_ref, settled_ref = O.distance_map(maze)
    assert stats["tree_used"] == ROOM
    assert (r.dist == d_ref).all() and (after == g_ref).all() and r.max == max_ref and r.settled == settled_ref


@pytest.mark.gpu
def test_gpu_room_and_wave_agree(lab, O, monkeypatch):
    maze = lab.build_maze("arena", 1535, 1023, seed=1)
    start = (1000, 17)
    with lab.Grid.from_host(maze) as g:
        r1 = g.distance_map(start)
        a1 = g.download()
        assert g.stats()["tree_used"] == ROOM
    monkeypatch.setenv("LAB_TREE", "0")  # neither tree pipeline nor room: the tile wave
    with lab.Grid.from_host(maze) as g:
        r0 = g.distance_map(start)
        a0 = g.download()
        assert g.stats()["tree_used"] == 0
    assert (r0.dist == r1.dist).all() and (a0 == a1).all() and r0.max == r1.max and r0.settled == r1.settled


@pytest.mark.gpu
@pytest.mark.parametrize("hole", [(700, 700), (1, 1), (1021, 1533)])
def test_gpu_arena_with_one_wall_is_not_a_room(lab, O, hole):
    maze = lab.build_maze("arena", 1023, 1535, seed=1)
    maze[hole] &= ~np.uint16(lab.PATH_BIT)
    with lab.Grid.from_host(maze) as g:
        r = g.distance_map()
        after = g.download()
        stats = g.stats()
    g_ref, d_ref, max_ref, settled_ref = O.distance_map(maze)
    assert stats["tree_used"] == 0
    assert (r.dist == d_ref).all() and (after == g_ref).all() and r.max == max_ref and r.settled == settled_ref


@pytest.mark.gpu
def test_gpu_games_in_a_room(lab, O):
    maze = lab.build_maze("arena", 1023, 2047, seed=1)
    placed, s, f = lab.place_hunt(maze, seed=0xB2000001)
    with lab.Grid.from_host(placed) as g:
        res = g.hunt(s, f)
        after = g.download()
        assert g.stats()["tree_used"] == ROOM
    g_ref, w_ref, p_ref = O.canon_hunt(placed, s, f)
    assert (after == g_ref).all() and res.winner == w_ref and paths_equal(res.paths, [p_ref])
    placed, s, fin = lab.place_gather(maze, seed=0xB2000003)
    with lab.Grid.from_host(placed) as g:
        res = g.gather(s, fin)
        after = g.download()
        assert g.stats()["tree_used"] == ROOM
    g_ref, c_ref, p_ref = O.canon_gather(placed, s, fin)
    assert (after == g_ref).all() and (res.claimed_by == c_ref).all() and paths_equal(res.paths, p_ref)
