"""The oracle and the product's host-side code against vectors produced by the REAL reference
(tests/golden/*.npz, written by oracle/make_golden.py while /root/reference was present).
Runs anywhere: no reference, no GPU."""
import numpy as np
import pytest

from _util import GOLDEN, golden_id, load_golden, paths_equal


@pytest.mark.parametrize("path", GOLDEN, ids=golden_id)
def test_product_builders_match_reference_grids(lab, path):
    z, builder, rows, cols, seed, mod = load_golden(path)
    assert (lab.build_maze(builder, rows, cols, seed, mod) == z["grid"]).all()


@pytest.mark.parametrize("path", GOLDEN, ids=golden_id)
def test_oracle_distance_matches_reference(O, path):
    z, *_ = load_golden(path)
    g, d, mx, settled = O.distance_map(z["grid"])
    assert (d == z["dist"]).all() and (g == z["dist_grid"]).all()
    assert mx == int(z["dist_max"][0]) and settled == int((z["dist"] != O.INF).sum())
    # the un-hooked stdout of the reference painter: whichever channel it picked at random, the
    # restated colour arithmetic (distance.cc:55-62) reproduces every printed triple
    rgb = z["dist_rgb"].astype(np.int32)
    assert any((O.painter_rgb(d, mx, ch) == rgb).all() for ch in range(3))
    assert ((rgb[..., 0] == -1) == (z["dist"] == O.INF)).all()


@pytest.mark.parametrize("path", GOLDEN, ids=golden_id)
def test_oracle_hunter_gatherer_match_reference(O, path):
    z, *_ = load_golden(path)
    for gname, game in (("hunt", 0), ("gather", 1)):
        placed, s = z[f"{gname}_placed"], z[f"{gname}_start"]
        for oi, order in enumerate([(0, 1, 2, 3), (3, 1, 0, 2)]):
            g, paths, win = O.solver_sequential(game, placed, [s] * 4, order)
            assert (g == z[f"{gname}_o{oi}_grid"]).all()
            assert win == int(z[f"{gname}_o{oi}_winner"][0])
            assert paths_equal(paths, [z[f"{gname}_o{oi}_path{k}"] for k in range(4)])


@pytest.mark.parametrize("path", GOLDEN, ids=golden_id)
def test_placement_matches_reference(O, lab, path):
    z, builder, rows, cols, seed, mod = load_golden(path)
    grid = z["grid"]
    for i, s in enumerate(range(seed, seed + 6)):
        assert O.pick_random_point(grid, s) == tuple(z["pick"][i])
        assert lab.pick_random_point(grid, s) == tuple(z["pick"][i])
    assert (O.set_corner_starts(grid) == z["corner_starts"]).all()
    assert (lab.set_corner_starts(grid) == z["corner_starts"]).all()
    # hunt / gather placement (the generator used the oracle's; the product's must agree, and both
    # must agree with where the reference's own games put the bits for the same seed)
    for name, place_o, place_l in (("hunt", O.place_hunt, lab.place_hunt), ("gather", O.place_gather, lab.place_gather)):
        go, so, fo = place_o(grid, seed + 100)
        gl, sl, fl = place_l(grid, seed + 100)
        assert (go == gl).all() and (go == z[f"{name}_placed"]).all()
        assert tuple(so) == tuple(sl) and (np.asarray(fo) == np.asarray(fl)).all()
    keep = np.uint16(O.START_BIT | O.FINISH_BIT | O.PATH_BIT)
    for name, place in (("hunt", lab.place_hunt), ("gather", lab.place_gather), ("corners", lab.place_corners)):
        placed = place(grid, seed + 200)[0]
        assert ((placed & keep) == z[f"game_{name}_start_finish_bits"]).all(), name
