"""Developer tool: time lab_distance_paint (the GPU heat-map painter) on a resident distance map.
usage: python tools/time_paint.py BUILDER N [REPS]"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import __graft_entry__ as graft
lab = graft.load_package()
builder, n, reps = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]) if len(sys.argv) > 3 else 10
maze = lab.build_maze(builder, n, n, seed=0xB2000002)
stream = torch.cuda.current_stream()
with lab.Grid.from_host(maze) as g:
    g.set_stream(stream.cuda_stream)
    levels = torch.empty((n, n), dtype=torch.int32, device="cuda")
    pixels = torch.empty((n, n), dtype=torch.int32, device="cuda")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    g.bind_levels(0, levels.data_ptr())
    g.bind_pixels(pixels.data_ptr())
    r = g.distance_map(want_dist=False)
    ms = []
    for i in range(reps + 2):
        flush.fill_(1)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        g.paint(i % 3, want_pixels=False)
        b.record(stream)
        torch.cuda.synchronize()
        if i >= 2:
            ms.append(a.elapsed_time(b))
    cells = n * n
    t = sorted(ms)[len(ms) // 2]
    print(json.dumps({"case": f"paint {builder} {n}", "max": r.max, "ms_median": t, "GB_per_s": cells * 8 / t / 1e6,
                      "frac_of_6545.9": cells * 8 / t / 1e6 / 6545.9, "bytes_per_square": 8}))
