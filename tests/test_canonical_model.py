"""The canonical (lock-step) games the GPU implements, checked against the restated reference
threads run under a sequential schedule -- no reference build needed.

north_star: solver outputs must be valid wall-respecting paths of exactly the reference BFS
shortest length, with identical gather/corners finish sets."""
import numpy as np
import pytest

from _util import valid_shortest_path

CASES = [("rdfs", 51, 111), ("kruskal", 65, 131), ("wilson", 63, 65), ("grid", 127, 129), ("arena", 33, 47)]


@pytest.mark.parametrize("builder,rows,cols", CASES)
def test_hunt_same_length_and_bracketed_paint(lab, O, builder, rows, cols):
    maze = lab.build_maze(builder, rows, cols, seed=11)
    for seed in range(5):
        placed, s, f = lab.place_hunt(maze, seed)
        g_seq, paths_seq, win_seq = O.solver_sequential(O.HUNT, placed, [s] * 4, (0, 1, 2, 3))
        g_can, win_can, path_can = O.canon_hunt(placed, s, f)
        assert len(path_can) == len(paths_seq[win_seq])  # exactly the reference BFS length
        assert valid_shortest_path(placed, path_can, s, f, len(path_can))
        d = O.bfs_levels(placed, *s)
        T = int(d[f])
        seq_paint = (g_seq & O.PAINT_MASK) != 0
        can_paint = (g_can & O.PAINT_MASK) != 0
        assert (can_paint == (d < T)).all()
        assert not seq_paint[d > T].any() and seq_paint[d < T].all()


@pytest.mark.parametrize("builder,rows,cols", CASES)
def test_gather_same_finish_sets(lab, O, builder, rows, cols):
    maze = lab.build_maze(builder, rows, cols, seed=12)
    for seed in range(5):
        placed, s, fin = lab.place_gather(maze, seed)
        g_seq, paths_seq, _ = O.solver_sequential(O.GATHER, placed, [s] * 4, (0, 1, 2, 3))
        g_can, claimed, paths_can = O.canon_gather(placed, s, fin)
        d = O.bfs_levels(placed, *s)
        # same multiset of path lengths = the four finish levels; every finish claimed in both
        assert sorted(len(p) for p in paths_can) == sorted(len(p) for p in paths_seq) == sorted(int(d[r, c]) for r, c in fin)
        assert sorted(claimed.tolist()) == [0, 1, 2, 3]
        for r, c in fin:
            assert g_seq[r, c] & O.CACHE_MASK and g_can[r, c] & O.CACHE_MASK
        for k in range(4):
            j = int(np.where(claimed == k)[0][0])
            assert valid_shortest_path(placed, paths_can[k], s, fin[j], int(d[fin[j][0], fin[j][1]]))


@pytest.mark.parametrize("builder,rows,cols", CASES)
def test_corners_winner_is_nearest_corner(lab, O, builder, rows, cols):
    maze = lab.build_maze(builder, rows, cols, seed=13)
    for seed in range(3):
        placed, starts, f = lab.place_corners(maze, seed)
        g_can, w, path = O.canon_corners(placed, starts, f)
        ds = [int(O.bfs_levels(placed, int(r), int(c))[f]) for r, c in starts]
        assert w == int(np.argmin(ds)) and len(path) == min(ds)
        assert valid_shortest_path(placed, path, starts[w], f, min(ds))
        # a sequential schedule of the reference's hunters finds the same length for its winner
        _, paths_seq, win_seq = O.solver_sequential(O.HUNT, placed, starts, (w, 0, 1, 2, 3)[:4] if w == 0 else (w,) + tuple(k for k in range(4) if k != w))
        assert win_seq == w and len(paths_seq[w]) == min(ds)
