"""Generate tests/golden/*.npz from the REAL reference (oracle/_ref, built from /root/reference).

Run in the authoring container only (the reference does not exist on the GPU box):
    python oracle/make_golden.py
The vectors pin, without the reference being present:
  * the builders  (grid bytes for builder x size x seed x mod),
  * painters/distance.cc:129-153 (levels, max, measure bits; painter colours of the un-hooked stdout),
  * solvers/bfs_threads.cc hunter / gatherer under sequential schedules (grids, paths, winner),
  * solvers/solve_utilities.cc point placement (pick_random_point, set_corner_starts),
  * the start/finish squares the full games (Bfs::hunt/gather/corners, real threads) place for a seed.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[0] = ROOT  # (the script directory would shadow the `oracle` package with oracle.py)
from oracle import oracle as O  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
CASES = [  # builder, rows, cols, seed, mod
    ("arena", 9, 9, 1, None), ("arena", 33, 47, 2, None),
    ("rdfs", 51, 111, 0xB2000001, None), ("rdfs", 65, 65, 3, "cross"),
    ("grid", 51, 111, 4, None), ("grid", 127, 129, 5, "cross"), ("grid", 63, 65, 6, "x"),
    ("kruskal", 51, 111, 7, None), ("kruskal", 129, 127, 0xB2000002, None), ("kruskal", 65, 131, 8, "x"),
    ("wilson", 51, 111, 9, None), ("wilson", 127, 65, 0xB2000005, "cross"),
]


def main():
    assert O.have_ref(), "build oracle/_ref first (make -C oracle ref)"
    os.makedirs(OUT, exist_ok=True)
    for builder, rows, cols, seed, mod in CASES:
        grid = O.ref_build(builder, rows, cols, seed, mod)
        rec = {"grid": grid, "meta": np.array([rows, cols, seed & 0xFFFFFFFF, {None: 0, "cross": 1, "x": 2}[mod]], np.int64)}
        # distance
        g_after, dist, mx, rgb = O.ref_distance(grid, seed=seed + 17)
        rec["dist_grid"] = g_after
        rec["dist"] = np.where(dist == np.uint64(2**64 - 1), np.uint64(O.INF), dist).astype(np.uint32)
        rec["dist_max"] = np.array([mx], np.int64)
        rec["dist_rgb"] = rgb.astype(np.int16)
        # placement
        rec["pick"] = np.array([O.ref_pick_random_point(grid, s) for s in range(seed, seed + 6)], np.int32)
        rec["corner_starts"] = O.ref_set_corner_starts(grid).astype(np.int32)
        # hunter / gatherer white-box, sequential schedules
        for gname, game in (("hunt", 0), ("gather", 1)):
            placed, s, f = (O.place_hunt if game == 0 else O.place_gather)(grid, seed + 100)
            rec[f"{gname}_placed"] = placed
            rec[f"{gname}_start"] = np.asarray(s, np.int32)
            rec[f"{gname}_finish"] = np.asarray(f, np.int32)
            for oi, order in enumerate([(0, 1, 2, 3), (3, 1, 0, 2)]):
                g2, paths, win = O.ref_solver_sequential(game, placed, [s] * 4, order)
                rec[f"{gname}_o{oi}_grid"] = g2
                rec[f"{gname}_o{oi}_winner"] = np.array([win], np.int32)
                for k in range(4):
                    rec[f"{gname}_o{oi}_path{k}"] = paths[k].astype(np.int32)
        # the full games with real threads: only the placement is schedule independent
        for gname, game in (("hunt", 0), ("gather", 1), ("corners", 2)):
            g3 = O.ref_game(game, grid, seed + 200)
            rec[f"game_{gname}_start_finish_bits"] = (g3 & np.uint16(O.START_BIT | O.FINISH_BIT | O.PATH_BIT)).astype(np.uint16)
        name = f"{builder}_{rows}x{cols}_s{seed & 0xFFFFFFFF:x}_{mod or 'none'}.npz"
        np.savez_compressed(os.path.join(OUT, name), **rec)
        print("wrote", name, os.path.getsize(os.path.join(OUT, name)), "bytes")


if __name__ == "__main__":
    main()
