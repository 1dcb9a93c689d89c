"""Developer tool: per-phase SM-cycle breakdown of the search kernel (needs the LAB_PROFILE build)."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("LAB_LIB", os.path.join(ROOT, "multithreading-with-mazes_b200", "liblabyrinth_b200_prof.so"))
import __graft_entry__ as graft
lab = graft.load_package()
cases = sys.argv[1:] or ["kruskal:4095", "wilson:4095", "arena:4095", "rdfs:2047", "grid:4095"]
for a in cases:
    builder, n = a.split(":")[0], int(a.split(":")[1])
    maze = lab.build_maze(builder, n, n, seed=0xB2000002)
    with lab.Grid.from_host(maze) as g:
        for rep in range(2):
            g.upload(maze)
            r = g.distance_map(want_dist=False)
        t, s = g.timing(), g.stats()
    act = max(1, s["activations"])
    print(json.dumps({
        "case": a, "search_ms": round(t["search_ms"], 3), "max_level": r.max, "activations": s["activations"], "empty": s["empty"],
        "levels": s["levels"], "cyc_per_level": round(s["cyc_loop"] / max(1, s["levels"]), 1),
        "load_cyc_per_act": round(s["cyc_load"] / act), "publish_cyc_per_act": round(s["cyc_publish"] / act),
        "loop_cyc_per_act": round(s["cyc_loop"] / act), "step_cyc_per_level": round(s["cyc_step"] / max(1, s["levels"]), 1),
        "emit_cyc_per_cell": round(s["cyc_emit"] / max(1, s["settled"]), 1), "emit_cyc_per_batch": round(s["cyc_emit"] / max(1, s["batches"])),
        "step_cyc_per_batch": round(s["cyc_step"] / max(1, s["batches"])), "batches": s["batches"], "cells_per_batch": round(s["settled"] / max(1, s["batches"]), 1), "idle_Mcyc": round(s["cyc_idle"] / 1e6, 1),
        "busy_Mcyc": round((s["cyc_load"] + s["cyc_loop"] + s["cyc_publish"]) / 1e6, 1),
        "kernel_Mcyc_x_warps": round(t["search_ms"] * 1e-3 * 1.9e9 * s["worker_warps"] / 1e6, 1),
        "pushes": s["queue_pushes"], "direct": s["direct"]}))
