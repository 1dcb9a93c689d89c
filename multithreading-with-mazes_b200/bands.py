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
    def tree_bind(self, tile_base: int, ntiles_global: int):
        """Allocates (once per maze size) the global crossing arrays and binds them.  Returns the dict of tensors."""
        torch = self.torch
        key = (tile_base, ntiles_global)
        if self._tree_bufs is None or self._tree_bufs["key"] != key:
            n = ntiles_global * 256
            dev = self.row_device
            b = {"key": key}
            for name in ("succ0", "s0", "s1", "r0", "r1"):
                b[name] = torch.empty(n + 2, dtype=torch.int32, device=dev)
            b["loc"] = torch.empty(n, dtype=torch.int32, device=dev)
            b["inmask"] = torch.zeros(ntiles_global * 4, dtype=torch.int64, device=dev)
            b["alist"] = torch.empty(n, dtype=torch.int32, device=dev)
            self._tree_bufs = b
        b = self._tree_bufs
        self._check(self._f("tree_bind")(self.h, c_uint32(tile_base), c_uint32(ntiles_global), self._ptr(b["succ0"]), self._ptr(b["s0"]),
                                         self._ptr(b["s1"]), self._ptr(b["r0"]), self._ptr(b["r1"]), self._ptr(b["loc"]),
                                         self._ptr(b["inmask"]), self._ptr(b["alist"])))
        return b

    def tree_count(self) -> Tuple[int, int, int]:
        v, e, x = c_uint64(), c_uint64(), c_uint32()
        self._check(self._f("tree_count")(self.h, byref(v), byref(e), byref(x)))
        return int(v.value), int(e.value), int(x.value)

    def tree_walk(self, squares_global: int, pairs_global: int, alist_base: int):
        self._check(self._f("tree_walk")(self.h, c_uint64(squares_global), c_uint64(pairs_global), c_uint32(alist_base)))

    def tree_rank(self, arcs_global: int):
        self._check(self._f("tree_rank")(self.h, c_uint32(arcs_global)))

    def tree_flood(self) -> Tuple[int, int, int]:
        u, st, err = c_uint32(), c_uint64(), c_uint32()
        self._check(self._f("tree_flood")(self.h, byref(u), byref(st), byref(err)))
        return int(u.value), int(st.value), int(err.value)

    def tree_levels(self, settled_global: int) -> bool:
        used = ctypes.c_int32()
        self._check(self._f("tree_levels")(self.h, c_uint64(settled_global), byref(used)))
        return used.value == 1

    def tree_abort(self):
        self._check(self._f("tree_abort")(self.h))


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
    engine.begin(src_local)
    top, bottom = engine.path_rows()
    above_bottom, below_top = _exchange(dist, rank, world, top, bottom, engine, group)
    engine.set_neighbour_paths(above_bottom, below_top)
    flag = torch.zeros(1, dtype=torch.int32, device=engine.row_device)
    rounds = 0
    tree = use_tree and getattr(engine, "supports_tree", False) and _sharded_tree(engine, start_global, row0, rows_local, group)
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
    mx, settled = engine.finish()
    agg = torch.tensor([mx, settled], dtype=torch.int64, device=engine.row_device)
    every = [torch.empty_like(agg) for _ in range(world)]
    dist.all_gather(every, agg, group=group)
    every = [e.tolist() for e in every]
    return {"max": max(int(e[0]) for e in every), "settled": sum(int(e[1]) for e in every), "rounds": rounds, "tree": bool(tree)}


def _share(dist, t, offs, lens, rank, world, group):
    """Every rank contributes t[offs[rank] : offs[rank] + lens[rank]]; afterwards all ranks hold all slices."""
    if world == 1:
        return
    flat = dist.get_backend(group) == "nccl"
    if flat and len(set(lens)) == 1 and all(offs[r] == offs[0] + r * lens[0] for r in range(world)):
        dist.all_gather_into_tensor(t[offs[0] : offs[0] + world * lens[0]], t[offs[rank] : offs[rank] + lens[rank]].clone(), group=group)
        return
    if flat:  # uneven pieces: one all-gather of padded pieces, then the valid parts into place
        longest = max(lens)
        mine = t.new_empty(longest)
        mine[: lens[rank]].copy_(t[offs[rank] : offs[rank] + lens[rank]])
        every = t.new_empty(world * longest)
        dist.all_gather_into_tensor(every, mine, group=group)
        for r in range(world):
            if r != rank and lens[r]:
                t[offs[r] : offs[r] + lens[r]].copy_(every[r * longest : r * longest + lens[r]])
        return
    for r in range(world):  # (gloo) one broadcast per band
        if lens[r]:
            dist.broadcast(t[offs[r] : offs[r] + lens[r]], src=dist.get_global_rank(group, r) if group is not None else r, group=group)


def _sharded_tree(engine, start_global, row0, rows_local, group) -> bool:
    """The spanning-tree pipeline across the bands.  Collective.  True: the bands' level fields are final (only
    engine.finish() is left); False: the maze is not one tree, the fields are as engine.begin() left them."""
    import torch
    import torch.distributed as dist

    rank = dist.get_rank(group)
    world = dist.get_world_size(group)
    dev = engine.row_device
    tiles_x = engine.cols_padded // TS

    def gather(values):
        """every rank's list of ints -> list (by rank) of lists, on the host"""
        t = torch.tensor(values, dtype=torch.int64, device=dev)
        if world == 1:
            return [[int(x) for x in t.tolist()]]
        if dist.get_backend(group) == "nccl":
            out = torch.empty(world * len(values), dtype=torch.int64, device=dev)
            dist.all_gather_into_tensor(out, t, group=group)
            return [[int(x) for x in row] for row in out.view(world, len(values)).tolist()]
        every = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(every, t, group=group)
        return [[int(x) for x in e.tolist()] for e in every]

    # who holds which tiles (bands are whole tile rows, in rank order); cached with the arrays
    lay = getattr(engine, "_tree_layout", None)
    if lay is None or lay["mine"] != (row0, rows_local, world):
        split = [tuple(x) for x in gather([row0, rows_local])]
        tile_rows = [(n + TS - 1) // TS for _, n in split]
        bases = [(r0 // TS) * tiles_x for r0, _ in split]
        ntiles = [tr * tiles_x for tr in tile_rows]
        total = sum(ntiles)
        ok = not (any(r0 % TS for r0, _ in split) or any(bases[r] != sum(ntiles[:r]) for r in range(world)) or total * 256 + 2 >= 1 << 32)
        lay = {"mine": (row0, rows_local, world), "split": split, "bases": bases, "ntiles": ntiles, "total": total, "ok": ok}
        engine._tree_layout = lay
    if not lay["ok"]:
        return False
    split, bases, ntiles, total = lay["split"], lay["bases"], lay["ntiles"], lay["total"]
    b = engine.tree_bind(bases[rank], total)
    nslots = total * 256

    counts = gather(list(engine.tree_count()))
    squares, pairs = sum(c[0] for c in counts), sum(c[1] for c in counts)
    if squares < 2 or pairs + 1 != squares:
        return False  # (nothing has been written yet)
    arcs = [c[2] for c in counts]
    arc_offs = [sum(arcs[:r]) for r in range(world)]
    engine.tree_walk(squares, pairs, arc_offs[rank])
    # the tour's segments of every band, its two ends from the band that holds the start, the list of crossings
    _share(dist, b["succ0"], [x * 256 for x in bases], [x * 256 for x in ntiles], rank, world, group)
    owner = next((r for r, (r0, n) in enumerate(split) if r0 <= start_global[0] < r0 + n), 0)
    if world > 1:
        dist.broadcast(b["succ0"][nslots : nslots + 2], src=dist.get_global_rank(group, owner) if group is not None else owner, group=group)
    _share(dist, b["alist"], arc_offs, arcs, rank, world, group)
    engine.tree_rank(sum(arcs))
    flags = gather(list(engine.tree_flood()))
    settled = sum(f[1] for f in flags)
    if any(f[0] or f[2] for f in flags) or settled != squares:
        engine.tree_abort()
        return False
    _share(dist, b["inmask"], [x * 4 for x in bases], [x * 4 for x in ntiles], rank, world, group)
    _share(dist, b["loc"], [x * 256 for x in bases], [x * 256 for x in ntiles], rank, world, group)
    return engine.tree_levels(settled)
