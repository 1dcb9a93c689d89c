"""Row-band sharding of the distance map across GPUs.

One process per GPU (``torch.distributed``, NCCL over NVLink/NVSwitch); rank r holds a contiguous
band of rows of the maze.  The search itself is local (the persistent tile kernel runs each band to
its local fixpoint); between two local searches neighbouring ranks exchange ONE ROW of levels
each way (cols x 4 bytes) and all ranks agree, with a one-word all-reduce, whether any border tile
was re-activated.  The loop ends when nobody was.  Levels are global BFS levels, so the result is
bit-identical to the single-GPU search.

Perfect mazes (one tree: every reference builder without ``-m``) do not take that loop -- their diameter
crosses the band borders thousands of times -- but the spanning-tree pipeline of csrc/device/tree.cuh, sharded:
walk, flood and fix-up run on each band's own tiles, the two pointer-jumping phases run over all the maze's
border crossings on every rank, and between the phases the ranks exchange their slices of the (global) crossing
arrays: four all-gathers and a few one-word all-reduces per map, whatever the diameter
(``_sharded_tree`` below; protocol in include/labyrinth_b200.h).

``torch`` is plumbing here (device buffers for the exchanged rows, the process group); every
kernel is the library's own.
"""
from __future__ import annotations

import ctypes
from ctypes import byref, c_int, c_uint32, c_uint64, c_void_p
from typing import List, Optional, Tuple

TS = 64
INF = 0xFFFFFFFF


class _Trace:
    """LAB_BANDS_TRACE=1: wall time of every phase of a sharded map (device synchronised after each, so the phases
    do not overlap as they do otherwise); rank 0 prints one line per map on stderr.  Developer aid."""

    def __init__(self, device):
        import os

        self.on = os.environ.get("LAB_BANDS_TRACE", "0") not in ("", "0")
        self.cuda = self.on and getattr(device, "type", "cpu") == "cuda"
        self.marks = []
        if self.on:
            self.mark("start")

    def mark(self, name):
        if not self.on:
            return
        import time

        if self.cuda:
            import torch

            torch.cuda.synchronize()
        self.marks.append((name, time.perf_counter()))

    def report(self, rank):
        if self.on and rank == 0:
            import sys

            parts = [f"{b[0]} {1e6 * (b[1] - a[1]):.0f}us" for a, b in zip(self.marks, self.marks[1:])]
            print("[lab bands] " + " | ".join(parts) + f" | total {1e6 * (self.marks[-1][1] - self.marks[0][1]):.0f}us", file=sys.stderr, flush=True)


def split_rows(rows: int, world: int) -> List[Tuple[int, int]]:
    """Band r = rows [row0, row0 + n).  Every band but the last is a multiple of 64 rows."""
    tiles = (rows + TS - 1) // TS
    out = []
    for r in range(world):
        r0 = min(rows, (r * tiles // world) * TS)
        r1 = min(rows, ((r + 1) * tiles // world) * TS)
        out.append((r0, r1 - r0))
    if any(n <= 0 for _, n in out):
        raise ValueError(f"{rows} rows are too few for {world} bands of whole tiles")
    return out


class _BandEngine:
    """One band: ctypes over <prefix>band_* entry points that share one protocol (include/labyrinth_b200.h).
    Subclasses set self.L (the library), self.h (the band's handle), self.prefix, self.row_device."""

    supports_tree = True

    def _setup_rows(self):
        torch = self.torch
        self.cols_padded = int(self._f("cols_padded")(self.h))
        dev = self.row_device
        self._u8 = [torch.empty(self.cols_padded, dtype=torch.uint8, device=dev) for _ in range(2)]
        self._u32 = [torch.empty(self.cols_padded, dtype=torch.int32, device=dev) for _ in range(2)]
        self._tree_bufs = None

    def _f(self, name):
        return getattr(self.L, self.prefix + name)

    def _check(self, rc):
        if rc != 0:
            raise RuntimeError(f"{self.prefix}* call failed: {rc}")

    @staticmethod
    def _ptr(t):
        return c_void_p(t.data_ptr()) if t is not None else None

    def begin(self, src_local: Optional[Tuple[int, int]]):
        r, c = src_local if src_local is not None else (-1, -1)
        self._check(self._f("begin")(self.h, c_int(r), c_int(c)))

    def path_rows(self):
        self._check(self._f("path_rows")(self.h, self._ptr(self._u8[0]), self._ptr(self._u8[1])))
        return self._u8[0], self._u8[1]

    def set_neighbour_paths(self, above_bottom, below_top):
        self._check(self._f("set_neighbour_paths")(self.h, self._ptr(above_bottom), self._ptr(below_top)))

    def relax(self):
        self._check(self._f("relax")(self.h))

    def export_edges(self):
        self._check(self._f("export_edges")(self.h, self._ptr(self._u32[0]), self._ptr(self._u32[1])))
        return self._u32[0], self._u32[1]

    def import_edges(self, above_bottom, below_top) -> int:
        n = c_uint32()
        self._check(self._f("import_edges")(self.h, self._ptr(above_bottom), self._ptr(below_top), byref(n)))
        return int(n.value)

    def finish(self) -> Tuple[int, int]:
        mx, st = c_uint64(), c_uint64()
        self._check(self._f("finish")(self.h, byref(mx), byref(st)))
        return int(mx.value), int(st.value)

    def new_row(self, like):
        return self.torch.empty_like(like)

    # -- the spanning-tree pipeline across bands (lab_band_tree_*) --
    def tree_bind(self, lay, rank: int):
        """Allocates (once per layout) the crossing arrays, the short list over the bands' border tile rows and the
        exchange chunks, and binds them.  lay: {"bases", "ntiles", "total", "tiles_x"}."""
        torch = self.torch
        key = (tuple(lay["bases"]), tuple(lay["ntiles"]), rank)
        if self._tree_bufs is None or self._tree_bufs["key"] != key:
            n = lay["total"] * 256
            dev = self.row_device
            world = len(lay["bases"])
            b = {"key": key}
            for name in ("succ0", "s0", "s1", "r0", "r1"):
                b[name] = torch.empty(n + 2, dtype=torch.int32, device=dev)
            b["loc"] = torch.empty(n, dtype=torch.int32, device=dev)
            b["inmask"] = torch.zeros(lay["total"] * 4, dtype=torch.int64, device=dev)
            b["alist"] = torch.empty(lay["ntiles"][rank] * 256, dtype=torch.int32, device=dev)
            rows = []  # every band's first and last tile row (once, if it has only one)
            for base, nt in zip(lay["bases"], lay["ntiles"]):
                rows.append(base)
                if nt > lay["tiles_x"]:
                    rows.append(base + nt - lay["tiles_x"])
            slots = torch.cat([torch.arange(t * 256, (t + lay["tiles_x"]) * 256, dtype=torch.int64) for t in rows])
            b["border_list"] = slots.to(torch.int32).to(dev)
            b["layout"] = torch.tensor([[x, y] for x, y in zip(lay["bases"], lay["ntiles"])], dtype=torch.int32, device=dev)
            for stage in (1, 2, 3):
                words = int(self._f("tree_chunk_words")(self.h, c_int(stage)))
                b[f"chunk{stage}"] = torch.zeros(words, dtype=torch.int32, device=dev)
                b[f"all{stage}"] = torch.zeros(world * words, dtype=torch.int32, device=dev)
            self._tree_bufs = b
        b = self._tree_bufs
        if b.get("bound"):
            return b
        b["bound"] = True
        self._check(self._f("tree_bind")(self.h, c_uint32(lay["bases"][rank]), c_uint32(lay["total"]), self._ptr(b["succ0"]), self._ptr(b["s0"]),
                                         self._ptr(b["s1"]), self._ptr(b["r0"]), self._ptr(b["r1"]), self._ptr(b["loc"]),
                                         self._ptr(b["inmask"]), self._ptr(b["alist"]), self._ptr(b["border_list"]),
                                         c_uint32(b["border_list"].numel()), self._ptr(b["layout"]), c_int(len(lay["bases"])), c_int(rank)))
        return b

    def tree_count(self):
        """-> squares, pairs, crossings, then the bounding box r0, r1, c0, c1 of the band's passable squares"""
        v, e, x = c_uint64(), c_uint64(), c_uint32()
        box = (ctypes.c_int32 * 4)()
        self._check(self._f("tree_count")(self.h, byref(v), byref(e), byref(x), box))
        return (int(v.value), int(e.value), int(x.value)) + tuple(int(b) for b in box)

    def room(self, r0: int, r1: int, c0: int, c1: int, src_r: int, src_c: int):
        self._check(self._f("room")(self.h, c_int(r0), c_int(r1), c_int(c0), c_int(c1), c_int(src_r), c_int(src_c)))

    def tree_rank_local(self, squares_global: int, pairs_global: int, chunk1):
        self._check(self._f("tree_rank_local")(self.h, c_uint64(squares_global), c_uint64(pairs_global), self._ptr(chunk1)))

    def tree_flood(self, all1, root_rank: int, chunk2):
        self._check(self._f("tree_flood")(self.h, self._ptr(all1), c_int(root_rank), self._ptr(chunk2)))

    def tree_prefix_local(self, all2, chunk3):
        self._check(self._f("tree_prefix_local")(self.h, self._ptr(all2), self._ptr(chunk3)))

    def tree_levels(self, all3) -> bool:
        used = ctypes.c_int32()
        self._check(self._f("tree_levels")(self.h, self._ptr(all3), byref(used)))
        return used.value == 1


class GpuBandEngine(_BandEngine):
    """One band on one GPU: the lab_band_* entry points of liblabyrinth_b200; buffers are CUDA torch tensors."""

    prefix = "lab_band_"

    def __init__(self, lab, squares, device=None):
        import torch

        self.lab, self.torch = lab, torch
        self.squares = squares  # int16 CUDA tensor [rows_local, cols], owned by the caller
        rows, cols = squares.shape
        self.grid = lab.Grid.from_device_ptr(rows, cols, squares.data_ptr(), keepalive=squares)
        self.grid.set_stream(torch.cuda.current_stream().cuda_stream)
        self.L = lab.lib()
        self.h = self.grid._h
        self.row_device = squares.device
        self._setup_rows()

    def _check(self, rc):
        if rc != 0:
            raise self.lab.LabError(rc, self.L.lab_last_error().decode(errors="replace"))


def _exchange(dist, rank, world, up_payload, down_payload, engine, group):
    """Send my first row up and my last row down; receive the neighbours' facing rows."""
    ops, from_up, from_down = [], None, None
    if rank > 0:
        from_up = engine.new_row(up_payload)
        ops += [dist.P2POp(dist.isend, up_payload, rank - 1, group), dist.P2POp(dist.irecv, from_up, rank - 1, group)]
    if rank < world - 1:
        from_down = engine.new_row(down_payload)
        ops += [dist.P2POp(dist.isend, down_payload, rank + 1, group), dist.P2POp(dist.irecv, from_down, rank + 1, group)]
    if ops:
        for req in dist.batch_isend_irecv(ops):
            req.wait()
    return from_up, from_down


def sharded_distance_map(engine, start_global: Tuple[int, int], row0: int, rows_local: int, group=None, max_rounds: int = 1 << 20,
                         use_tree: bool = True):
    """Run the sharded distance map on this rank's band.  Collective: every rank of `group` calls it.

    Returns {"max", "settled", "rounds"} (global values; the level field stays in the engine)."""
    import torch
    import torch.distributed as dist

    rank = dist.get_rank(group)
    world = dist.get_world_size(group)
    sr, sc = start_global
    src_local = (sr - row0, sc) if row0 <= sr < row0 + rows_local else None
    tr = _Trace(engine.row_device)
    engine.begin(src_local)
    tr.mark("begin")
    top, bottom = engine.path_rows()
    if world > 1 and dist.get_backend(group) == "nccl":  # one flat all-gather beats two send/recv pairs for rows this small
        n = top.numel()
        every = torch.empty(world * 2 * n, dtype=top.dtype, device=top.device)
        dist.all_gather_into_tensor(every, torch.cat([top, bottom]), group=group)
        every = every.view(world, 2, n)
        above_bottom = every[rank - 1, 1] if rank > 0 else None
        below_top = every[rank + 1, 0] if rank < world - 1 else None
    else:
        above_bottom, below_top = _exchange(dist, rank, world, top, bottom, engine, group)
    tr.mark("path rows")
    engine.set_neighbour_paths(above_bottom, below_top)
    tr.mark("cross")
    flag = torch.zeros(1, dtype=torch.int32, device=engine.row_device)
    rounds = 0
    tree = use_tree and getattr(engine, "supports_tree", False) and _sharded_tree(engine, start_global, row0, rows_local, group, tr)
    while not tree:
        rounds += 1
        engine.relax()
        top, bottom = engine.export_edges()
        above_bottom, below_top = _exchange(dist, rank, world, top, bottom, engine, group)
        activated = engine.import_edges(above_bottom, below_top)
        flag.fill_(1 if activated else 0)
        dist.all_reduce(flag, op=dist.ReduceOp.MAX, group=group)
        if int(flag.item()) == 0 or rounds >= max_rounds:
            break
    tr.mark("search")
    mx, settled = engine.finish()
    tr.mark("finish")
    agg = torch.tensor([mx, settled], dtype=torch.int64, device=engine.row_device)
    out = torch.empty(2 * world, dtype=torch.int64, device=engine.row_device)
    _all_gather(dist, out, agg, world, group)
    every = out.view(world, 2).tolist()
    tr.mark("results")
    tr.report(rank)
    return {"max": max(int(e[0]) for e in every), "settled": sum(int(e[1]) for e in every), "rounds": rounds, "tree": tree is True, "road": "wave" if not tree else ("tree" if tree is True else "room")}


def _all_gather(dist, out, chunk, world, group):
    """out (world x len(chunk)) <- every rank's chunk"""
    if world == 1:
        out.copy_(chunk)
    elif dist.get_backend(group) == "nccl":
        dist.all_gather_into_tensor(out, chunk, group=group)
    else:
        dist.all_gather(list(out.view(world, -1).unbind(0)), chunk, group=group)


def tree_layout(split, tiles_x: int):
    """Who holds which tiles: bands are whole tile rows, in rank order.  None if they are not."""
    ntiles = [((n + TS - 1) // TS) * tiles_x for _, n in split]
    bases = [(r0 // TS) * tiles_x for r0, _ in split]
    total = sum(ntiles)
    if any(r0 % TS for r0, _ in split) or any(bases[r] != sum(ntiles[:r]) for r in range(len(split))) or total * 256 + 2 >= 1 << 32:
        return None
    return {"split": list(split), "bases": bases, "ntiles": ntiles, "total": total, "tiles_x": tiles_x}


def _sharded_tree(engine, start_global, row0, rows_local, group, tr=None) -> bool:
    """The spanning-tree pipeline across the bands.  Collective.  True: the bands' level fields are final (only
    engine.finish() is left); False: the maze is not one tree, the fields are as engine.begin() left them."""
    import torch
    import torch.distributed as dist

    rank = dist.get_rank(group)
    world = dist.get_world_size(group)
    dev = engine.row_device

    def gather(values):
        """every rank's list of ints -> list (by rank) of lists, on the host"""
        t = torch.tensor(values, dtype=torch.int64, device=dev)
        out = torch.empty(world * len(values), dtype=torch.int64, device=dev)
        _all_gather(dist, out, t, world, group)
        return [[int(x) for x in row] for row in out.view(world, len(values)).tolist()]

    cached = getattr(engine, "_tree_layout", None)
    if cached is None or cached[0] != (row0, rows_local, world):
        cached = ((row0, rows_local, world), tree_layout([tuple(x) for x in gather([row0, rows_local])], engine.cols_padded // TS))
        engine._tree_layout = cached
    lay = cached[1]
    if lay is None:
        return False
    mark = tr.mark if tr is not None else (lambda name: None)
    b = engine.tree_bind(lay, rank)
    mark("bind")
    counts = engine.tree_count()
    mark("count")
    counts = gather(list(counts))
    mark("gather counts")
    squares, pairs = sum(c[0] for c in counts), sum(c[1] for c in counts)
    if squares < 2 or pairs + 1 != squares:
        # not a tree (nothing has been written yet) -- but maybe an open room: the passable squares fill their bounding box
        boxes = [(split_r0 + c[3], split_r0 + c[4], c[5], c[6]) for c, (split_r0, _) in zip(counts, lay["split"]) if c[0]]
        if boxes:
            R0, R1 = min(b[0] for b in boxes), max(b[1] for b in boxes)
            C0, C1 = min(b[2] for b in boxes), max(b[3] for b in boxes)
            if squares == (R1 - R0 + 1) * (C1 - C0 + 1):
                engine.room(max(R0, row0) - row0, min(R1, row0 + rows_local - 1) - row0, C0, C1, start_global[0] - row0, start_global[1])
                mark("room")
                return "room"
        return False
    owner = next((r for r, (r0, n) in enumerate(lay["split"]) if r0 <= start_global[0] < r0 + n), 0)
    engine.tree_rank_local(squares, pairs, b["chunk1"])
    mark("walk + rank")
    _all_gather(dist, b["all1"], b["chunk1"], world, group)
    mark("all-gather 1")
    engine.tree_flood(b["all1"], owner, b["chunk2"])
    mark("short rank + flood")
    _all_gather(dist, b["all2"], b["chunk2"], world, group)
    mark("all-gather 2")
    engine.tree_prefix_local(b["all2"], b["chunk3"])
    mark("parent + prefix")
    _all_gather(dist, b["all3"], b["chunk3"], world, group)
    mark("all-gather 3")
    used = engine.tree_levels(b["all3"])
    mark("short prefix + fix-up")
    return used
