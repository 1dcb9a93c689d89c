"""Developer tool: run one search a few times (target command for ncu captures)."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as graft
lab = graft.load_package()
what, builder, n, reps = sys.argv[1], sys.argv[2], int(sys.argv[3]), int(sys.argv[4]) if len(sys.argv) > 4 else 2
maze = lab.cached_maze(builder, n, n, seed=0xB2000002)
if what == "gather":
    maze, start, fin = lab.place_gather(maze, seed=0xB2000003)
if what == "hunt":
    maze, start, fin = lab.place_hunt(maze, seed=0xB2000003)
with lab.Grid.from_host(maze) as g:
    for rep in range(reps):
        g.upload(maze)
        if what == "distance":
            r = g.distance_map(want_dist=False)
            out = {"max": r.max, "settled": r.settled}
        elif what == "hunt":
            r = g.hunt(start, fin, want_path=False)
            out = {"path_len": r.path_len, "painted": r.painted}
        else:
            r = g.gather(start, fin, want_paths=False)
            out = {"path_len": r.path_len, "painted": r.painted}
    print(json.dumps({"case": f"{what} {builder} {n}", **out, "ms": g.timing(), "stats": g.stats()}))
