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
ap.add_argument("--game", default="distance", choices=["distance", "hunt", "gather", "corners"])
ap.add_argument("--mod", default=None, choices=[None, "cross", "x"], help="-m modification of the maze (loops: the tile wave across bands)")
ap.add_argument("--super", action="store_true", help="bands of whole super-tile rows (the lattice road across bands)")
ap.add_argument("--check-property", action="store_true", help="verify every band on the device through the BFS equations (tests/test_properties.py), "
                "with one halo row of levels from the neighbouring bands: for sizes beyond the oracle's reach")
ap.add_argument("--seed", type=lambda x: int(x, 0), default=0xB2000005)
ap.add_argument("--cached", action="store_true", help="build the maze once (rank 0), cache it as a raw maze file, map it on the other ranks")
a = ap.parse_args()
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
lab = graft.load_package()
from mazes_b200 import bands
if a.cached:
    if rank == 0:
        lab.cached_maze(a.builder, a.rows, a.cols, seed=a.seed, mmap=True)
    dist.barrier()
    maze = lab.cached_maze(a.builder, a.rows, a.cols, seed=a.seed, mmap=True)
else:
    maze = lab.build_maze(a.builder, a.rows, a.cols, seed=a.seed, mod=a.mod)
if a.game != "distance":
    # the sharded solvers: place the game on the host as the reference does, run it band by band, compare with the oracle
    if a.game == "hunt":
        # (--super: the lattice road across bands wants the start on a cell -- take the first seed that puts it on one)
        seed = next(q for q in range(0xB2000006, 0xB2000006 + 200) if not a.super or all(int(x) % 2 == 1 for x in lab.place_hunt(maze, seed=q)[1]))
        placed, s, f = lab.place_hunt(maze, seed=seed)
    elif a.game == "corners":
        placed, s, f = lab.place_corners(maze, seed=0xB2000008)
    else:
        placed, s, f = lab.place_gather(maze, seed=0xB2000007)
    r0, n = (bands.split_rows_super if a.super else bands.split_rows)(a.rows, world)[rank]
    for rep in range(a.reps):
        sq = torch.from_numpy(placed[r0:r0 + n].view(np.int16).copy()).cuda()
        eng = bands.GpuBandEngine(lab, sq)
        torch.cuda.synchronize(); dist.barrier()
        t0 = time.perf_counter()
        if a.game == "hunt":
            res = bands.sharded_hunt(eng, tuple(int(x) for x in s), tuple(int(x) for x in f), r0, n)
        elif a.game == "corners":
            res = bands.sharded_corners(eng, [tuple(int(x) for x in q) for q in s], tuple(int(x) for x in f), r0, n)
            res["path_len"] = res["level"]
        else:
            res = bands.sharded_gather(eng, tuple(int(x) for x in s), [tuple(int(x) for x in q) for q in f], r0, n)
        torch.cuda.synchronize()
        t = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.MAX)
        if rank == 0:
            print(json.dumps({"game": a.game, "builder": a.builder, "rows": a.rows, "cols": a.cols, "world": world, "rep": rep,
                              "wall_ms": round(float(t.item()) * 1e3, 3), "path_len": res["path_len"], "painted": res["painted"], "road": res["road"],
                              "rounds": res.get("rounds")}), flush=True)
    if a.check:
        O = graft.load_oracle()
        if a.game == "corners":
            g_ref, w_ref, p_ref = O.canon_corners(placed, s, f)
            ok = bool((sq.cpu().numpy().view(np.uint16) == g_ref[r0:r0 + n]).all() and res["winner"] == w_ref and res["level"] == len(p_ref))
        elif a.game == "hunt":
            g_ref, w_ref, p_ref = O.canon_hunt(placed, s, f)
            ok = bool((sq.cpu().numpy().view(np.uint16) == g_ref[r0:r0 + n]).all() and res["winner"] == w_ref and (res["path"] == np.asarray(p_ref)).all())
        else:
            g_ref, c_ref, p_ref = O.canon_gather(placed, s, f)
            got = sq.cpu().numpy().view(np.uint16)
            print(f"rank {rank}: grid diffs {int((got != g_ref[r0:r0 + n]).sum())} claimed {list(res['claimed_by'])} vs {list(c_ref)} paths "
                  f"{[bool(res['paths'][k].shape == np.asarray(p_ref[k]).reshape(-1, 2).shape and (res['paths'][k] == np.asarray(p_ref[k]).reshape(-1, 2)).all()) for k in range(4)]}", flush=True)
            if (got != g_ref[r0:r0 + n]).any():
                rr, cc = np.nonzero(got != g_ref[r0:r0 + n])
                print(f"rank {rank}: first diffs {[(int(a) + r0, int(b), hex(int(got[a, b])), hex(int(g_ref[a + r0, b]))) for a, b in list(zip(rr, cc))[:6]]}", flush=True)
            ok = bool((sq.cpu().numpy().view(np.uint16) == g_ref[r0:r0 + n]).all() and (np.asarray(res["claimed_by"]) == c_ref).all()
                      and all((res["paths"][k] == np.asarray(p_ref[k]).reshape(-1, 2)).all() for k in range(4)))
        print(f"rank {rank}: band rows [{r0},{r0 + n}) {'OK' if ok else 'MISMATCH'}", flush=True)
    dist.destroy_process_group()
    sys.exit(0)
r0, n = (bands.split_rows_super if a.super else bands.split_rows)(a.rows, world)[rank]
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
if a.check_property:
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from test_properties import check_field
    # one halo row of levels and Squares from each neighbour, then the BFS equations on the band's own rows
    up_lv, dn_lv = torch.full((1, a.cols), -1, dtype=torch.int32, device="cuda"), torch.full((1, a.cols), -1, dtype=torch.int32, device="cuda")
    # (NCCL has no 16-bit integer type: the Squares' rows travel as bytes)
    up_sq, dn_sq = torch.zeros((1, a.cols), dtype=torch.int16, device="cuda"), torch.zeros((1, a.cols), dtype=torch.int16, device="cuda")
    sq_first, sq_last = sq[:1].contiguous().view(torch.uint8), sq[-1:].contiguous().view(torch.uint8)
    ops = []
    if rank > 0:
        ops += [dist.P2POp(dist.isend, lv[:1].contiguous(), rank - 1), dist.P2POp(dist.irecv, up_lv, rank - 1),
                dist.P2POp(dist.isend, sq_first, rank - 1), dist.P2POp(dist.irecv, up_sq.view(torch.uint8), rank - 1)]
    if rank < world - 1:
        ops += [dist.P2POp(dist.isend, lv[-1:].contiguous(), rank + 1), dist.P2POp(dist.irecv, dn_lv, rank + 1),
                dist.P2POp(dist.isend, sq_last, rank + 1), dist.P2POp(dist.irecv, dn_sq.view(torch.uint8), rank + 1)]
    if ops:
        for req in dist.batch_isend_irecv(ops):
            req.wait()
    first, last = (1 if rank > 0 else 0), (1 if rank > 0 else 0) + n
    ext_lv = torch.cat(([up_lv] if rank > 0 else []) + [lv] + ([dn_lv] if rank < world - 1 else []))
    ext_sq = torch.cat(([up_sq] if rank > 0 else []) + [sq] + ([dn_sq] if rank < world - 1 else []))
    s_ext = (start[0] - r0 + first, start[1]) if r0 <= start[0] < r0 + n else (-10, -10)
    try:
        reached, mx_band = check_field(torch, ext_lv, ext_sq, s_ext, lab, first_row=first, last_row=last)
        tot = torch.tensor([reached], dtype=torch.int64, device="cuda"); dist.all_reduce(tot)
        mxt = torch.tensor([mx_band], dtype=torch.int64, device="cuda"); dist.all_reduce(mxt, op=dist.ReduceOp.MAX)
        ok = int(tot.item()) == res["settled"] and int(mxt.item()) == res["max"]
        print(f"rank {rank}: band rows [{r0},{r0 + n}) BFS equations {'OK' if ok else 'COUNT/MAX MISMATCH'} (reached {reached}, band max {mx_band})", flush=True)
    except AssertionError as e:
        print(f"rank {rank}: band rows [{r0},{r0 + n}) BFS equations VIOLATED: {e}", flush=True)
if a.check:
    O = graft.load_oracle()
    g_ref, d_ref, mx, settled = O.distance_map(maze)
    ok = bool((lv.cpu().numpy().view(np.uint32) == d_ref[r0:r0 + n]).all() and (sq.cpu().numpy().view(np.uint16) == g_ref[r0:r0 + n]).all()
              and res["max"] == mx and res["settled"] == settled)
    print(f"rank {rank}: band rows [{r0},{r0 + n}) {'OK' if ok else 'MISMATCH'}", flush=True)
dist.destroy_process_group()
