"""Developer tool: bfs-gather WITH paths on a 16383^2 Kruskal maze, the paths from the tour's ranks (LAB_LAT_PATH=1) and walked (0)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as graft
lab = graft.load_package()
maze = lab.cached_maze("kruskal", 16383, 16383, seed=0xB2000002)
maze, start, fin = lab.place_gather(maze, seed=0xB2000003)
for path_on in ("1", "0"):
    os.environ["LAB_LAT_PATH"] = path_on
    with lab.Grid.from_host(maze) as g:
        for rep in range(3):
            g.upload(maze)
            r = g.gather(start, fin, want_paths=True)
        print("LAB_LAT_PATH", path_on, "path_len", list(r.path_len), {k: round(v, 3) for k, v in g.timing().items()}, flush=True)
