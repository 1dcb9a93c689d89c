"""Sharded distance map over row bands, one rank per GPU (launch with torchrun).  Verifies every
band against the oracle when --check is given (developer tool, not the bench)."""
import argparse, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import torch.distributed as dist
import __graft_entry__ as graft

ap = argparse.ArgumentParser()
ap.add_argument("--builder", default="kruskal"); ap.add_argument("--rows", type=int, default=4095); ap.add_argument("--cols", type=int, default=4095)
ap.add_argument("--check", action="store_true"); ap.add_argument("--reps", type=int, default=3)
a = ap.parse_args()
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
lab = graft.load_package()
from mazes_b200 import bands
maze = lab.build_maze(a.builder, a.rows, a.cols, seed=0xB2000005)
r0, n = bands.split_rows(a.rows, world)[rank]
pristine = torch.from_numpy(maze[r0:r0 + n].view(np.int16).copy()).cuda()
sq = torch.empty_like(pristine)
lv = torch.empty((n, a.cols), dtype=torch.int32, device="cuda")
eng = bands.GpuBandEngine(lab, sq)
eng.grid.bind_levels(0, lv.data_ptr())
start = lab.distance_start(a.rows, a.cols)
for rep in range(a.reps):
    sq.copy_(pristine)
    torch.cuda.synchronize(); dist.barrier()
    t0 = time.perf_counter()
    res = bands.sharded_distance_map(eng, start, r0, n)
    torch.cuda.synchronize()
    el = time.perf_counter() - t0
    t = torch.tensor([el], dtype=torch.float64, device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(json.dumps({"builder": a.builder, "rows": a.rows, "cols": a.cols, "world": world, "rep": rep, "wall_ms": round(float(t.item()) * 1e3, 3),
                          "Gcells_per_s": round(res["settled"] / float(t.item()) / 1e9, 3), **res}), flush=True)
if a.check:
    O = graft.load_oracle()
    g_ref, d_ref, mx, settled = O.distance_map(maze)
    ok = bool((lv.cpu().numpy().view(np.uint32) == d_ref[r0:r0 + n]).all() and (sq.cpu().numpy().view(np.uint16) == g_ref[r0:r0 + n]).all()
              and res["max"] == mx and res["settled"] == settled)
    print(f"rank {rank}: band rows [{r0},{r0 + n}) {'OK' if ok else 'MISMATCH'}", flush=True)
dist.destroy_process_group()
