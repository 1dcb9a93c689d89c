"""Randomised differential tests: random builder / size / modification / start alignment, the three single-field entry
points against the oracle.  CPU: the kernel bodies on the warp emulator (random warp counts and schedules); GPU: the C ABI."""
import numpy as np
import pytest

from _util import Emu, paths_equal


def random_case(lab, rng, max_half):
    b = str(rng.choice(["arena", "kruskal", "wilson", "rdfs", "grid"]))
    rows, cols = int(rng.integers(3, max_half)) * 2 + 1, int(rng.integers(3, max_half)) * 2 + 1
    mod = str(rng.choice(["cross", "x"])) if b in ("grid", "kruskal") and rng.random() < 0.4 else None
    maze = lab.build_maze(b, rows, cols, seed=int(rng.integers(1 << 30)), mod=mod)
    if b == "arena" and rng.random() < 0.3:  # an arena that is no longer an open room
        maze[int(rng.integers(1, rows - 1)), int(rng.integers(1, cols - 1))] &= ~np.uint16(lab.PATH_BIT)
    return b, mod, maze


def test_emulated_random_cases(lab, O):
    emu = Emu()
    rng = np.random.default_rng(2026)
    for it in range(14):
        b, mod, maze = random_case(lab, rng, 70)
        tag = (it, b, mod, maze.shape)
        g_ref, d_ref, mx_ref, st_ref = O.distance_map(maze)
        g, d, mx, st, _ = emu.distance(maze, nwarps=int(rng.integers(1, 6)), sched=int(rng.integers(0, 50)), misalign=int(rng.integers(0, 8)))
        assert (d == d_ref).all() and (g == g_ref).all() and mx == mx_ref and st == st_ref, tag
        placed, s, f = lab.place_hunt(maze, seed=int(rng.integers(1 << 30)))
        g_ref, w_ref, p_ref = O.canon_hunt(placed, s, f)
        g, w, _, paths, _ = emu.solver(1, placed, [s], [f], int(rng.integers(1, 5)), int(rng.integers(0, 50)))
        assert (g == g_ref).all() and w == w_ref and paths_equal([paths[0]], [p_ref]), tag
        placed, s, fin = lab.place_gather(maze, seed=int(rng.integers(1 << 30)))
        g_ref, c_ref, p_ref = O.canon_gather(placed, s, fin)
        g, _, c, paths, _ = emu.solver(2, placed, [s], fin, int(rng.integers(1, 5)), int(rng.integers(0, 50)))
        assert (g == g_ref).all() and (c == c_ref).all() and paths_equal(paths, p_ref), tag


@pytest.mark.gpu
def test_gpu_random_cases(lab, O):
    rng = np.random.default_rng(4052)
    for it in range(120):
        b, mod, maze = random_case(lab, rng, 330)
        tag = (it, b, mod, maze.shape)
        g_ref, d_ref, mx_ref, st_ref = O.distance_map(maze)
        with lab.Grid.from_host(maze) as g:
            r = g.distance_map()
            after = g.download()
            px = g.paint(int(rng.integers(0, 3)))
        assert (r.dist == d_ref).all() and (after == g_ref).all() and r.max == mx_ref and r.settled == st_ref, tag
        assert ((px >> 24) != 0).sum() == st_ref, tag
        placed, s, f = lab.place_hunt(maze, seed=int(rng.integers(1 << 30)))
        g_ref, w_ref, p_ref = O.canon_hunt(placed, s, f)
        with lab.Grid.from_host(placed) as g:
            res = g.hunt(s, f)
            after = g.download()
        assert (after == g_ref).all() and res.winner == w_ref and paths_equal(res.paths, [p_ref]), tag
        placed, s, fin = lab.place_gather(maze, seed=int(rng.integers(1 << 30)))
        g_ref, c_ref, p_ref = O.canon_gather(placed, s, fin)
        with lab.Grid.from_host(placed) as g:
            res = g.gather(s, fin)
            after = g.download()
        assert (after == g_ref).all() and (res.claimed_by == c_ref).all() and paths_equal(res.paths, p_ref), tag
        placed, st4, f = lab.place_corners(maze, seed=int(rng.integers(1 << 30)))
        g_ref, w_ref, p_ref = O.canon_corners(placed, st4, f)
        with lab.Grid.from_host(placed) as g:
            res = g.corners(st4, f)
            after = g.download()
        assert (after == g_ref).all() and res.winner == w_ref, tag
