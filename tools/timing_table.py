"""Developer tool: device-side timings and work counters for a few mazes (not the bench)."""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import __graft_entry__ as graft

lab = graft.load_package()
cases = [("kruskal", 1023, None), ("kruskal", 4095, None), ("wilson", 4095, None), ("arena", 4095, None),
         ("grid", 4095, "cross"), ("rdfs", 2047, None), ("arena", 16383, None)]
if len(sys.argv) > 1:
    cases = [(a.split(":")[0], int(a.split(":")[1]), None) for a in sys.argv[1:]]
for builder, n, mod in cases:
    t0 = time.time()
    maze = lab.build_maze(builder, n, n, seed=0xB2000002, mod=mod)
    tb = time.time() - t0
    with lab.Grid.from_host(maze) as g:
        for rep in range(3):
            g.upload(maze)
            r = g.distance_map(want_dist=False)
            t = g.timing(); s = g.stats()
        cells = n * n
        print(json.dumps({"case": f"distance {builder}{'+' + mod if mod else ''} {n}x{n}", "build_s": round(tb, 2), "max": r.max,
                          "settled": r.settled, "ms": {k: round(v, 3) for k, v in t.items()},
                          "Gcells_settled_per_s": round(r.settled / t["total_ms"] / 1e6, 3),
                          "roofline_frac": round(8.0 * cells / (t["total_ms"] * 1e-3) / 6545.9e9, 4), "stats": s}))
        if builder == "arena":
            placed, st, fin = lab.place_gather(maze, seed=0xB2000003)
            for rep in range(2):
                g.upload(placed)
                res = g.gather(st, fin, want_paths=False)
                t = g.timing(); s = g.stats()
            print(json.dumps({"case": f"gather {builder} {n}x{n}", "path_len": res.path_len, "painted": res.painted,
                              "ms": {k: round(v, 3) for k, v in t.items()},
                              "roofline_frac": round(4.0 * cells / (t["total_ms"] * 1e-3) / 6545.9e9, 4), "stats": s}))
