"""Parity at sizes the CPU oracle cannot reach in seconds, through a size-independent property.

A level field L over a maze is THE breadth-first field of painters/distance.cc:129-153 iff
    L(start) = 0,   L(x) = 1 + min over the passable 4-neighbours y of L(y)   for every other reached square,
and walls / unreached squares carry no level: following a neighbour one level lower from any square gives a path of
exactly L(x) steps to the start (so L >= distance), and L(x) <= L(y) + 1 along a shortest path gives L <= distance.
The check is a handful of torch ops on the device, row block by row block, so it runs at BASELINE's 32767 x 32767 --
on every road a search can take: spanning-tree pipeline (Kruskal, Wilson), open room (arena), tile wave (grid + cross,
arena with a pillar).  The Square grid's side effects are checked the same way (measure bit <=> reached)."""
import numpy as np
import pytest

INF = 0xFFFFFFFF


def check_field(torch, levels, squares, start, lab, block=2048, first_row=0, last_row=None):
    """levels: int32 CUDA tensor (bit pattern of uint32), squares: int16 CUDA tensor (after the map), start: (r, c).
    first_row / last_row: check these rows only (a row band with one halo row of its neighbours above and below:
    tools/run_bands.py --check-property)."""
    rows, cols = levels.shape
    last_row = rows if last_row is None else last_row
    big = 1 << 40
    reached_total, max_level = 0, 0
    for r0 in range(first_row, last_row, block):
        r1 = min(last_row, r0 + block)
        lo, hi = max(0, r0 - 1), min(rows, r1 + 1)
        L = levels[lo:hi].to(torch.int64) & 0xFFFFFFFF
        L = torch.where(L == INF, torch.full_like(L, big), L)
        # one ring of "no level" around the block, except where the maze really continues (the rows above / below)
        P = torch.nn.functional.pad(L, (1, 1, 1 if lo == r0 else 0, 1 if hi == r1 else 0), value=big)
        n = r1 - r0
        C = P[1 : n + 1, 1:-1]
        nmin = torch.minimum(torch.minimum(P[0:n, 1:-1], P[2 : n + 2, 1:-1]), torch.minimum(P[1 : n + 1, :-2], P[1 : n + 1, 2:]))
        sq = squares[r0:r1].to(torch.int32) & 0xFFFF
        passable = (sq & lab.PATH_BIT) != 0
        reached = C < big
        is_start = torch.zeros_like(reached)
        if r0 <= start[0] < r1:
            is_start[start[0] - r0, start[1]] = True
            assert int(C[start[0] - r0, start[1]]) == 0
        assert bool((reached == (passable | is_start)).all()), "a level on a wall, or a passable square without one"
        assert bool(torch.where(reached & ~is_start, C == nmin + 1, torch.ones_like(reached)).all()), "not 1 + min of the neighbours"
        assert bool((((sq & lab.MEASURE_BIT) != 0) == reached).all()), "measure bit != reached"
        reached_total += int(reached.sum())
        if bool(reached.any()):
            max_level = max(max_level, int(C[reached].max()))
    return reached_total, max_level


CASES = [("kruskal", 8191, None, 3), ("wilson", 8191, None, 3), ("arena", 32767, None, 2), ("arena", 8191, "pillar", 0), ("grid", 4095, "cross", 0)]


@pytest.mark.gpu
@pytest.mark.parametrize("builder,n,mod,road", CASES)
def test_gpu_levels_satisfy_the_bfs_equations_at_scale(lab, builder, n, mod, road):
    import torch

    maze = lab.build_maze(builder, n, n, seed=0xB2000002, mod=None if mod == "pillar" else mod)
    if mod == "pillar":
        maze[n // 4, n // 4] &= ~np.uint16(lab.PATH_BIT)
    start = lab.distance_start(n, n)
    squares = torch.from_numpy(maze.view(np.int16)).cuda()
    levels = torch.empty((n, n), dtype=torch.int32, device="cuda")
    g = lab.Grid.from_device_ptr(n, n, squares.data_ptr(), keepalive=squares)
    try:
        g.set_stream(torch.cuda.current_stream().cuda_stream)
        g.bind_levels(0, levels.data_ptr())
        res = g.distance_map(start, want_dist=False)
        assert g.stats()["tree_used"] == road
        torch.cuda.synchronize()
        reached, mx = check_field(torch, levels, squares, start, lab)
        assert reached == res.settled and mx == res.max
        passable = int(((maze & lab.PATH_BIT) != 0).sum()) + (0 if maze[start] & lab.PATH_BIT else 1)
        assert reached == passable  # (the reference's builders make connected mazes)
    finally:
        g.close()


@pytest.mark.parametrize("builder,mod", [("kruskal", None), ("arena", None), ("grid", "cross")])
def test_the_checker_itself(lab, O, builder, mod):
    """CPU: the property accepts the oracle's fields (blocks smaller than the maze: the seams are checked too) and
    rejects a field with one level off by one, a missing level and a missing measure bit."""
    import torch

    maze = lab.build_maze(builder, 131, 197, seed=3, mod=mod)
    g, d, mx, settled = O.distance_map(maze)
    start = lab.distance_start(131, 197)
    as_t = lambda a, t: torch.from_numpy(a.view(t).copy())
    assert check_field(torch, as_t(d, np.int32), as_t(g, np.int16), start, lab, block=50) == (settled, mx)
    ys, xs = np.nonzero(d != INF)
    for what in ("level", "hole", "bit"):
        d2, g2 = d.copy(), g.copy()
        y, x = ys[len(ys) // 2], xs[len(xs) // 2]
        if what == "level":
            d2[y, x] += 1
        elif what == "hole":
            d2[y, x] = INF
        else:
            g2[y, x] &= ~np.uint16(lab.MEASURE_BIT)
        with pytest.raises(AssertionError):
            check_field(torch, as_t(d2, np.int32), as_t(g2, np.int16), start, lab, block=50)


@pytest.mark.gpu
@pytest.mark.parametrize("builder,n", [("wilson", 8191), ("arena", 16383)])
def test_gpu_hunt_and_gather_invariants_at_scale(lab, builder, n):
    """SURVEY App. A.4's schedule-independent invariants of the games, at sizes beyond the oracle: the winner's path is a
    wall-respecting chain of exactly the BFS length from a neighbour of the finish to the start, repainted in the winner's
    colour only; every other square carries paint iff it is closer to the start than the finish is; gather: every finish
    claimed, colour k's paint == cache, and paint iff closer than the finish colour k claims."""
    import torch

    from _util import valid_shortest_path

    maze = lab.build_maze(builder, n, n, seed=0xB2000005)
    # levels from the start of the game (a field the test above proves exact at this size)
    placed, s, f = lab.place_hunt(maze, seed=0xB2000006)
    with lab.Grid.from_host(placed) as g:
        d = g.distance_map(tuple(s)).dist
    D = int(d[f[0], f[1]])
    with lab.Grid.from_host(placed) as g:
        res = g.hunt(s, f)
        after = g.download()
    assert res.winner == 0 and res.path_len[0] == D
    path = np.asarray(res.paths[0]).reshape(-1, 2)
    assert valid_shortest_path(placed, path, s, f, D)
    assert (d[path[:, 0], path[:, 1]] == np.arange(D - 1, -1, -1)).all()  # one level lower at every step
    paint = (after & lab.PAINT_MASK) >> 4
    want = np.where(d < D, 0xF, 0).astype(paint.dtype)
    want[path[:, 0], path[:, 1]] = 1  # the winner's (colour 0's) bit alone
    assert (paint == want).all() and ((after & ~np.uint16(lab.PAINT_MASK)) == (placed & ~np.uint16(lab.PAINT_MASK))).all()
    del after, paint, want
    # gather
    placed, s, fin = lab.place_gather(maze, seed=0xB2000007)
    with lab.Grid.from_host(placed) as g:
        d = g.distance_map(tuple(s)).dist
    with lab.Grid.from_host(placed) as g:
        res = g.gather(s, fin)
        after = g.download()
    Dk = sorted(int(d[r, c]) for r, c in fin)
    assert sorted(int(x) for x in res.claimed_by) == [0, 1, 2, 3]
    cache = ((after & lab.CACHE_MASK) >> 8).astype(np.int32)
    paint = ((after & lab.PAINT_MASK) >> 4).astype(np.int32)
    want = np.zeros_like(cache)
    for k in range(4):
        want |= np.where(d < Dk[k], 1 << k, 0)
    fronts = [np.asarray(p).reshape(-1, 2)[0] for p in res.paths if len(p)]
    differs = paint != want
    allowed = np.zeros_like(differs)
    for fr in fronts:
        allowed[fr[0], fr[1]] = True  # path.front() is recoloured (bfs_threads.cc:355-364)
    assert not (differs & ~allowed).any()
    want_cache = want.copy()
    for j, (r, c) in enumerate(fin):
        want_cache[r, c] |= 1 << int(res.claimed_by[j])  # the claim (bfs_threads.cc:158-161)
    assert (cache == want_cache).all()
    for k in range(4):
        j = int(np.nonzero(np.asarray(res.claimed_by) == k)[0][0])
        assert valid_shortest_path(placed, res.paths[k], s, fin[j], int(d[fin[j][0], fin[j][1]]))


@pytest.mark.gpu
@pytest.mark.parametrize("n", [8191])
def test_gpu_corners_invariants_at_scale(lab, n):
    """bfs-corners (solvers/bfs_threads.cc:420-453) beyond the oracle's sizes, on BASELINE configs[3]'s builder (grid, -m
    cross): four fields, one per corner start.  With d_k the exact level field from corner k (each proven a BFS field by
    the equations above, on the device), D = min_k d_k(finish): the winner is the lowest k with d_k(finish) == D, its
    path a wall-respecting chain of exactly D squares from a neighbour of the finish to its corner, and colour k paints
    exactly the squares with d_k < D; nothing else of a Square changes (static corners does not repaint the path)."""
    import torch

    from _util import valid_shortest_path

    maze = lab.build_maze("grid", n, n, seed=0xB2000002, mod="cross")
    placed, starts, f = lab.place_corners(maze, seed=0xB2000003)
    fields = []
    for k in range(4):
        sq = torch.from_numpy(placed.view(np.int16).copy()).cuda()
        lv = torch.empty((n, n), dtype=torch.int32, device="cuda")
        g = lab.Grid.from_device_ptr(n, n, sq.data_ptr(), keepalive=sq)
        try:
            g.set_stream(torch.cuda.current_stream().cuda_stream)
            g.bind_levels(0, lv.data_ptr())
            res = g.distance_map(tuple(int(x) for x in starts[k]), want_dist=False)
            torch.cuda.synchronize()
            reached, mx = check_field(torch, lv, sq, tuple(int(x) for x in starts[k]), lab)
            assert reached == res.settled and mx == res.max
            fields.append(lv.cpu().numpy().view(np.uint32))
        finally:
            g.close()
    Dk = [int(d[f[0], f[1]]) for d in fields]
    D = min(Dk)
    with lab.Grid.from_host(placed) as g:
        res = g.corners(starts, f)
        after = g.download()
    assert res.winner == Dk.index(D) and res.path_len[0] == D
    assert valid_shortest_path(placed, res.paths[0], tuple(int(x) for x in starts[res.winner]), f, D)
    path = np.asarray(res.paths[0]).reshape(-1, 2)
    assert (fields[res.winner][path[:, 0], path[:, 1]] == np.arange(D - 1, -1, -1)).all()
    want = np.zeros((n, n), np.uint16)
    for k in range(4):
        want |= np.where(fields[k] < D, np.uint16(1 << (4 + k)), np.uint16(0))
    assert ((after & lab.PAINT_MASK) == want).all()
    assert ((after & ~np.uint16(lab.PAINT_MASK)) == (placed & ~np.uint16(lab.PAINT_MASK))).all()
