import sys, json
sys.path.insert(0, ".")
import __graft_entry__ as graft
lab = graft.load_package()
maze = lab.cached_maze("kruskal", 16383, 16383, seed=0xB2000002)
maze, start, fin = lab.place_gather(maze, seed=0xB2000003)
for opt in (1, 0):
    with lab.Grid.from_host(maze) as g:
        g.set_option(lab.OPT_SOLVER_LEVELS, opt)
        for rep in range(4):
            g.upload(maze)
            r = g.gather(start, fin, want_paths=False)
        print("keep levels", opt, "painted", r.painted, {k: round(v, 3) for k, v in g.timing().items()})
