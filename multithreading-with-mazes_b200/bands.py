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


LAT_SUPER_TILES = 4  # LST of csrc/device/lattice.cuh: bands of the lattice road are whole super-tile rows


def split_rows_super(rows: int, world: int, super_tiles: int = LAT_SUPER_TILES) -> List[Tuple[int, int]]:
    """Bands of whole SUPER-TILE rows (4 x 64 rows), all of the same height but the last: what the lattice road across
    bands (``_sharded_lattice``) needs.  The other sharded roads take these bands too."""
    unit = super_tiles * TS
    sup_rows = (rows + unit - 1) // unit
    per = (sup_rows + world - 1) // world
    out = [(min(rows, r * per * unit), min(rows, (r + 1) * per * unit) - min(rows, r * per * unit)) for r in range(world)]
    if any(n <= 0 for _, n in out):
        raise ValueError(f"{rows} rows are too few for {world} bands of whole super-tile rows ({unit} rows each)")
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
        self._u32x4 = None  # corners: four fields' rows per buffer (allocated by begin_corners)
        self.nplanes = 1
        self._tree_bufs = None

    def _f(self, name):
        return getattr(self.L, self.prefix + name)

    def _check(self, rc):
        if rc != 0:
            raise RuntimeError(f"{self.prefix}* call failed: {rc}")

    @staticmethod
    def _ptr(t):
        return c_void_p(t.data_ptr()) if t is not None else None

    def begin(self, src_local: Optional[Tuple[int, int]], lattice_next: bool = False):
        r, c = src_local if src_local is not None else (-1, -1)
        self.nplanes = 1
        self._check(self._f("begin_lattice")(self.h, c_int(0), c_int(r), c_int(c), c_int(1 if lattice_next else 0)))

    def begin_corners(self, srcs_local):
        """Bfs::corners: four fields at once; srcs_local[k] = colour k's start in band rows, or None (another band's)."""
        flat = []
        for s_ in srcs_local:
            flat += list(s_) if s_ is not None else [-1, -1]
        self.nplanes = 4
        if self._u32x4 is None:
            self._u32x4 = [self.torch.empty(4 * self.cols_padded, dtype=self.torch.int32, device=self.row_device) for _ in range(2)]
        self._check(self._f("begin_corners")(self.h, (ctypes.c_int32 * 8)(*flat)))

    def path_rows(self):
        self._check(self._f("path_rows")(self.h, self._ptr(self._u8[0]), self._ptr(self._u8[1])))
        return self._u8[0], self._u8[1]

    def set_neighbour_paths(self, above_bottom, below_top):
        self._check(self._f("set_neighbour_paths")(self.h, self._ptr(above_bottom), self._ptr(below_top)))

    def relax(self):
        self._check(self._f("relax")(self.h))

    def export_edges(self):
        bufs = self._u32 if self.nplanes == 1 else self._u32x4  # (field k's row at [k * cols_padded, (k + 1) * cols_padded))
        self._check(self._f("export_edges")(self.h, self._ptr(bufs[0]), self._ptr(bufs[1])))
        return bufs[0], bufs[1]

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
            b["succ0"] = torch.empty(n + 2, dtype=torch.int32, device=dev)
            for name in ("sr0", "sr1"):  # (pointer, value) pairs
                b[name] = torch.empty(n + 2, dtype=torch.int64, device=dev)
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
        self._check(self._f("tree_bind")(self.h, c_uint32(lay["bases"][rank]), c_uint32(lay["total"]), self._ptr(b["succ0"]), self._ptr(b["sr0"]),
                                         self._ptr(b["sr1"]), self._ptr(b["loc"]),
                                         self._ptr(b["inmask"]), self._ptr(b["alist"]), self._ptr(b["border_list"]),
                                         c_uint32(b["border_list"].numel()), self._ptr(b["layout"]), c_int(len(lay["bases"])), c_int(rank)))
        return b

    def tree_count(self):
        """-> squares, pairs, crossings, then the bounding box r0, r1, c0, c1 of the band's passable squares"""
        v, e, x = c_uint64(), c_uint64(), c_uint32()
        box = (ctypes.c_int32 * 4)()
        self._check(self._f("tree_count")(self.h, byref(v), byref(e), byref(x), box))
        return (int(v.value), int(e.value), int(x.value)) + tuple(int(b) for b in box)

    # -- the solvers across bands (lab_band_begin_game, _level_at, _paint, _walk, _recolour, _finish_game) --
    def begin_game(self, game: int, src_local: Optional[Tuple[int, int]], lattice_next: bool = False):
        r, c = src_local if src_local is not None else (-1, -1)
        self.nplanes = 1
        self._check(self._f("begin_lattice")(self.h, c_int(game), c_int(r), c_int(c), c_int(1 if lattice_next else 0)))

    def level_at(self, r: int, c: int, plane: int = 0) -> int:
        v = c_uint32()
        self._check(self._f("level_at_plane")(self.h, c_int(plane), c_int(r), c_int(c), byref(v)))
        return int(v.value)

    def paint(self, rules) -> int:
        """rules: [(below, bits)] or [(below, bits, field)], at most 4 -> squares of this band that received a bit"""
        n = len(rules)
        below = (c_uint32 * 4)(*[r[0] for r in rules] + [0] * (4 - n))
        bits = (c_uint32 * 4)(*[r[1] for r in rules] + [0] * (4 - n))
        plane = (c_uint32 * 4)(*[(r[2] if len(r) > 2 else 0) for r in rules] + [0] * (4 - n))
        painted = c_uint64()
        self._check(self._f("paint_planes")(self.h, c_int(n), plane, below, bits, byref(painted)))
        return int(painted.value)

    def walk(self, r: int, c: int, include_first: bool, first_only: bool, repaint_bits: int, path_buf=None):
        """-> (squares walked, row, col where the walk stopped, level left); path_buf: int32 [cap, 2] tensor or None"""
        state = (ctypes.c_int64 * 4)()
        cap = int(path_buf.shape[0]) if path_buf is not None else 0
        self._check(self._f("walk")(self.h, c_int(r), c_int(c), c_int(1 if include_first else 0), c_int(1 if first_only else 0),
                                    c_uint32(repaint_bits), self._ptr(path_buf), c_uint64(cap), state))
        return int(state[0]), int(state[1]), int(state[2]), int(state[3])

    def recolour(self, r: int, c: int, paint_bits: int):
        self._check(self._f("recolour")(self.h, c_int(r), c_int(c), ctypes.c_uint16(paint_bits)))

    def finish_game(self):
        self._check(self._f("finish_game")(self.h))

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


    # -- the lattice road across bands (lab_band_lat_*) --
    def lat_bind(self, rank: int, per: int, world: int):
        """Whole-maze border list / weights / chunk arrays (allocated once per layout) and the band's place in them."""
        torch = self.torch
        key = (rank, per, world)
        b = getattr(self, "_lat_bufs", None)
        if b is None or b["key"] != key:
            tiles_x = self.cols_padded // TS
            n_b, n_w, n_c = c_uint64(), c_uint64(), c_uint64()
            self._check(self._f("lat_words")(c_int(tiles_x), c_int(per * world), byref(n_b), byref(n_w), byref(n_c)))
            dev = self.row_device
            b = {"key": key, "gb": torch.empty(int(n_b.value), dtype=torch.int64, device=dev),
                 "wsum": torch.zeros(int(n_w.value), dtype=torch.int32, device=dev), "chunk": torch.empty(int(n_c.value), dtype=torch.int32, device=dev),
                 "hdr": torch.zeros(8, dtype=torch.int32, device=dev), "all_hdr": torch.zeros(8 * world, dtype=torch.int32, device=dev),
                 "slice": int(n_b.value) // world, "reduce_words": int(n_w.value) - 1024}
            self._check(self._f("lat_bind")(self.h, c_int(rank * per), c_int(per), c_int(per * world), self._ptr(b["gb"]), self._ptr(b["wsum"]),
                                            self._ptr(b["chunk"])))
            self._lat_bufs = b
        return b

    def lat_link(self, hdr):
        self._check(self._f("lat_link")(self.h, self._ptr(hdr)))

    def lat_flood(self, all_hdr, world: int, zero_weights: bool):
        self._check(self._f("lat_flood")(self.h, self._ptr(all_hdr), c_int(world), c_int(1 if zero_weights else 0)))

    def lat_path_rank(self, r: int, c: int) -> Tuple[int, int]:
        """(rank of the entry arc of the component that holds band square (r, c), the square's level)"""
        rank, level = c_uint32(), c_uint32()
        self._check(self._f("lat_path_rank")(self.h, c_int(r), c_int(c), byref(rank), byref(level)))
        return int(rank.value), int(level.value)

    def lat_path_mark(self, rank: int, level: int, fin_local, start_local, repaint_bits: int, path_buf=None) -> int:
        """Marks this band's squares of the finish -> start path (no hand-over between bands) -> segments walked here"""
        fr, fc = fin_local if fin_local is not None else (-1, -1)
        sr, sc = start_local if start_local is not None else (-1, -1)
        n = c_uint64()
        cap = int(path_buf.shape[0]) if path_buf is not None else 0
        self._check(self._f("lat_path_mark")(self.h, c_uint32(rank), c_uint32(level), c_int(fr), c_int(fc), c_int(sr), c_int(sc), c_uint32(repaint_bits),
                                             self._ptr(path_buf), c_uint64(cap), byref(n)))
        return int(n.value)

    def lat_levels(self) -> bool:
        used = ctypes.c_int32()
        self._check(self._f("lat_levels")(self.h, byref(used)))
        return used.value == 1


class GpuBandEngine(_BandEngine):
    """One band on one GPU: the lab_band_* entry points of liblabyrinth_b200; buffers are CUDA torch tensors."""

    prefix = "lab_band_"
    supports_lattice = True

    def grid_cols(self):
        return self.grid.cols

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

    def or_bits(self, cells, bits):
        self.grid.or_bits(cells, bits)


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
                         use_tree: bool = True, use_lattice: bool = True):
    """Run the sharded distance map on this rank's band.  Collective: every rank of `group` calls it.

    Returns {"max", "settled", "rounds"} (global values; the level field stays in the engine)."""
    import torch
    import torch.distributed as dist

    rank = dist.get_rank(group)
    world = dist.get_world_size(group)
    sr, sc = start_global
    src_local = (sr - row0, sc) if row0 <= sr < row0 + rows_local else None
    tr = _Trace(engine.row_device)
    lattice_layout = _lattice_layout(engine, start_global, row0, rows_local, group) if (use_tree and use_lattice) else None
    if lattice_layout is not None:
        engine.begin(src_local, lattice_next=True)
    else:
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
    tree = False
    if use_tree and getattr(engine, "supports_tree", False):
        lattice = lattice_layout is not None and _sharded_lattice(engine, lattice_layout, group, tr)
        tree = "lattice" if lattice else _sharded_tree(engine, start_global, row0, rows_local, group, tr)
    if not tree:
        rounds = _wave_to_fixpoint(engine, group, max_rounds)
    tr.mark("search")
    mx, settled = engine.finish()
    tr.mark("finish")
    agg = torch.tensor([mx, settled], dtype=torch.int64, device=engine.row_device)
    out = torch.empty(2 * world, dtype=torch.int64, device=engine.row_device)
    _all_gather(dist, out, agg, world, group)
    every = out.view(world, 2).tolist()
    tr.mark("results")
    tr.report(rank)
    return {"max": max(int(e[0]) for e in every), "settled": sum(int(e[1]) for e in every), "rounds": rounds, "tree": tree is True or tree == "lattice",
            "road": "wave" if not tree else ("tree" if tree is True else tree)}


def _wave_to_fixpoint(engine, group, max_rounds: int = 1 << 20) -> int:
    """The tile wave, band by band (general mazes): local search to the band's fixpoint, one row of levels per open field
    to each neighbour, a one-word all-reduce; until no rank's border tiles were re-activated.  -> exchange rounds"""
    import torch
    import torch.distributed as dist

    rank, world = dist.get_rank(group), dist.get_world_size(group)
    flag = torch.zeros(1, dtype=torch.int32, device=engine.row_device)
    rounds = 0
    while True:
        rounds += 1
        engine.relax()
        top, bottom = engine.export_edges()
        above_bottom, below_top = _exchange(dist, rank, world, top, bottom, engine, group)
        activated = engine.import_edges(above_bottom, below_top)
        flag.fill_(1 if activated else 0)
        dist.all_reduce(flag, op=dist.ReduceOp.MAX, group=group)
        if int(flag.item()) == 0:
            return rounds
        if rounds >= max_rounds:
            raise RuntimeError(f"no fixpoint after {max_rounds} exchange rounds")


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


def _lattice_layout(engine, start_global, row0, rows_local, group):
    """(super-tile rows per band, rows of the whole maze) if the lattice road across bands applies -- bands of whole
    super-tile rows, all of one height but the last, odd sizes, the start on a cell -- else None.  Collective the first time."""
    import torch
    import torch.distributed as dist

    if not getattr(engine, "supports_lattice", False):
        return None
    world = dist.get_world_size(group)
    dev = engine.row_device
    cached = getattr(engine, "_lat_layout", None)
    if cached is None or cached[0] != (row0, rows_local, world):
        t = torch.tensor([row0, rows_local], dtype=torch.int64, device=dev)
        out = torch.empty(2 * world, dtype=torch.int64, device=dev)
        _all_gather(dist, out, t, world, group)
        split = [(int(a), int(b)) for a, b in out.view(world, 2).tolist()]
        unit = LAT_SUPER_TILES * TS
        per = (split[0][1] + unit - 1) // unit
        ok = all(r0 == r * per * unit for r, (r0, _) in enumerate(split)) and all(n == per * unit for _, n in split[:-1]) and 0 < split[-1][1] <= per * unit
        cached = ((row0, rows_local, world), (per, sum(n for _, n in split)) if ok else None)
        engine._lat_layout = cached
    if cached[1] is None:
        return None
    per, rows_global = cached[1]
    sr, sc = start_global
    if not (rows_global % 2 and int(engine.grid_cols()) % 2 and sr % 2 and sc % 2):
        return None
    return cached[1]


def _sharded_lattice(engine, layout, group, tr=None) -> bool:
    """The lattice road (csrc/device/lattice.cuh) across the bands.  Collective.  Two exchanges per map: the bands' slices
    of the border list (+ one header each) are all-gathered, every rank finishes the ranking over the whole list and
    floods its own tiles; the arcs' weights, scattered by tour position into a whole-maze array, are summed over the
    ranks; every rank scans them and writes its band's levels.  True: the bands' level fields are final; False: not
    applicable (bands not whole super-tile rows, even sizes, start not on a cell) or not one lattice tree -- nothing has
    been written and the general pipeline / the wave take over."""
    import torch
    import torch.distributed as dist

    rank = dist.get_rank(group)
    world = dist.get_world_size(group)
    mark = tr.mark if tr is not None else (lambda name: None)
    per, rows_global = layout
    b = engine.lat_bind(rank, per, world)
    engine.lat_link(b["hdr"])
    mark("lattice: link + local rank")
    _all_gather(dist, b["all_hdr"], b["hdr"], world, group)
    n = b["slice"]
    mine = b["gb"][rank * n:(rank + 1) * n]
    _all_gather(dist, b["gb"], mine if dist.get_backend(group) == "nccl" else mine.clone(), world, group)
    mark("lattice: all-gather border list")
    engine.lat_flood(b["all_hdr"], world, True)
    mark("lattice: rank + flood")
    if world > 1:
        # only the tour's positions (1 .. arcs of the whole maze) hold weights; the bands' counts sit in the array's tail
        arcs = int(b["all_hdr"].view(world, 8)[:, 4].sum().item())
        tail = b["reduce_words"] - 8
        dist.all_reduce(b["wsum"][: min(arcs + 2, tail)], op=dist.ReduceOp.SUM, group=group)
        dist.all_reduce(b["wsum"][tail: tail + 8], op=dist.ReduceOp.SUM, group=group)
    mark("lattice: all-reduce weights")
    used = engine.lat_levels()
    mark("lattice: scan + levels")
    return used


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


# --------------------------------------------------------------------------------------------------------------------
# The solvers across row bands (Bfs::hunt, Bfs::gather: solvers/bfs_threads.cc:243-280,333-369; canonical lock-step
# model), on mazes whose level field needs no wave: one tree (the sharded spanning-tree pipeline) or
# an open room.  The field is computed exactly as for the distance map; the games are then local passes (paint rules with
# thresholds every rank knows) plus the canonical paths, which are walked band by band: a walk that reaches its band's
# border is taken over by the neighbour (one tiny all-gather per hand-over); in a room every band writes its part of the
# two-segment path at once.

GAME_HUNT, GAME_GATHER = 1, 2
PAINT_SHIFT, CACHE_SHIFT, PAINT_MASK = 4, 8, 0x00F0


def _gather_ints(engine, values, group):
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    t = torch.tensor(values, dtype=torch.int64, device=engine.row_device)
    out = torch.empty(world * len(values), dtype=torch.int64, device=engine.row_device)
    _all_gather(dist, out, t, world, group)
    return [[int(x) for x in row] for row in out.view(world, len(values)).tolist()]


def _sharded_field(engine, game, start_global, row0, rows_local, group):
    """begin + the level field from `start_global`: 'tree' or 'room'.  Anything else is not sharded for the solvers yet."""
    import torch
    import torch.distributed as dist

    rank, world = dist.get_rank(group), dist.get_world_size(group)
    sr, sc = start_global
    tr = _Trace(engine.row_device)
    lattice_layout = _lattice_layout(engine, start_global, row0, rows_local, group)
    engine.begin_game(game, (sr - row0, sc) if row0 <= sr < row0 + rows_local else None, lattice_next=lattice_layout is not None)
    top, bottom = engine.path_rows()
    above_bottom, below_top = _exchange(dist, rank, world, top, bottom, engine, group)
    engine.set_neighbour_paths(above_bottom, below_top)
    road = False
    if lattice_layout is not None and _sharded_lattice(engine, lattice_layout, group, tr):
        return "lattice"  # (the same level field the general pipeline leaves; the tour's ranks stay for lat_path_*)
    if not road:
        road = _sharded_tree(engine, start_global, row0, rows_local, group, tr)
    if not road:
        # a general maze (loops: grid, -m cross, -m x): the tile wave across bands, to the field's fixpoint.  The canonical
        # model paints {d < D}: the whole field with the thresholds applied afterwards gives the same grid a bounded search does.
        _wave_to_fixpoint(engine, group)
        return "wave"
    return "tree" if road is True else "room"


def _band_split(engine, row0, rows_local, group):
    """[(first row, rows)] of every rank's band (all-gathered once per layout, kept on the engine)."""
    import torch.distributed as dist

    world = dist.get_world_size(group)
    cached = getattr(engine, "_split", None)
    if cached is None or cached[0] != (row0, rows_local, world):
        cached = ((row0, rows_local, world), [tuple(x) for x in _gather_ints(engine, [row0, rows_local], group)])
        engine._split = cached
    return cached[1]


def _level_of(engine, road, square, start_global, row0, rows_local, group):
    """BFS level of a global square, known to every rank afterwards (-1: unreachable)."""
    r, c = square
    if road == "room":  # (Manhattan distance: the closed form every rank can evaluate)
        return abs(r - start_global[0]) + abs(c - start_global[1])
    mine = engine.level_at(r - row0, c) if row0 <= r < row0 + rows_local else -1
    mine = -1 if mine == INF else mine
    return max(v[0] for v in _gather_ints(engine, [mine], group))


def _walk_path(engine, road, finish, start_global, level, repaint_bits, want_path, row0, rows_local, group):
    """The canonical path from `finish` (level `level`) to the start: repaints its squares (if repaint_bits) and returns it
    as an int32 [level, 2] array of global (row, col) pairs, finish side first (None unless want_path)."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    dev = engine.row_device
    full = torch.full((max(level, 1), 2), -1, dtype=torch.int32, device=dev) if want_path else None
    if level > 0:
        if road == "room":  # every band writes its own squares of the two segments, at their indices along the path
            engine.walk(finish[0] - row0, finish[1], False, False, repaint_bits, full)
            if want_path:
                full[:, 0] = torch.where(full[:, 0] >= 0, full[:, 0] + row0, full[:, 0])
        else:
            piece = torch.empty((level, 2), dtype=torch.int32, device=dev) if want_path else None
            cur, include_first, done = (finish[0], finish[1]), False, 0
            for _ in range(level + 2):  # (every hand-over makes progress: at most `level` of them)
                own = row0 <= cur[0] < row0 + rows_local
                state = [0, 0, 0, 0, 0]
                if own:
                    steps, r, c, left = engine.walk(cur[0] - row0, cur[1], include_first, False, repaint_bits, piece)
                    state = [1, steps, r + row0, c, left]
                    if want_path and steps:
                        full[done : done + steps, 0] = piece[:steps, 0] + row0
                        full[done : done + steps, 1] = piece[:steps, 1]
                got = next(v for v in _gather_ints(engine, state, group) if v[0] == 1)
                _, steps, r, c, left = got
                done += steps
                if left <= 0:
                    break
                # the walk stopped at its band's border: its next square is across it (one level lower, in the neighbour's band)
                owner_r0, owner_n = next((a, n) for a, n in _band_split(engine, row0, rows_local, group) if a <= r < a + n)
                cur = (r - 1, c) if r == owner_r0 else (r + 1, c)
                if not (r == owner_r0 or r == owner_r0 + owner_n - 1) or steps == 0 and include_first:
                    raise RuntimeError(f"path walk lost at ({r},{c}) with {left} levels to go")
                include_first = True
            else:
                raise RuntimeError("path walk did not reach the start")
    if not want_path:
        return None
    if world > 1:
        every = torch.empty((world,) + tuple(full.shape), dtype=torch.int32, device=dev)
        _all_gather(dist, every.view(-1), full.view(-1), world, group)
        full = every.max(dim=0).values
    return full[:level].cpu().numpy()


def _lattice_path(engine, finish, start_global, level, repaint_bits, want_path, row0, rows_local, group):
    """The finish -> start path on the lattice road, no hand-over: the band that holds the finish reads the rank of the entry
    arc of the finish's component (one all-gather of two ints), then every band marks the path's segments in its own rows
    from the whole-maze tour ranks it already holds, at their indices along the path."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    here = lambda sq: row0 <= sq[0] < row0 + rows_local
    mine = list(engine.lat_path_rank(finish[0] - row0, finish[1])) + [1] if here(finish) else [0, 0, 0]
    rank_f, level_f, _ = next(v for v in _gather_ints(engine, mine, group) if v[2] == 1)
    assert level_f == level
    full = torch.full((level, 2), -1, dtype=torch.int32, device=engine.row_device) if want_path else None
    engine.lat_path_mark(rank_f, level, (finish[0] - row0, finish[1]) if here(finish) else None,
                         (start_global[0] - row0, start_global[1]) if here(start_global) else None, repaint_bits, full)
    if not want_path:
        return None
    full[:, 0] = torch.where(full[:, 0] >= 0, full[:, 0] + row0, full[:, 0])
    if world > 1:
        every = torch.empty((world,) + tuple(full.shape), dtype=torch.int32, device=engine.row_device)
        _all_gather(dist, every.view(-1), full.view(-1), world, group)
        full = every.max(dim=0).values
    return full.cpu().numpy()


def sharded_hunt(engine, start_global, finish_global, row0: int, rows_local: int, group=None, want_path: bool = True):
    """Bfs::hunt on this rank's band (the caller placed start_bit / finish_bit).  Collective.
    -> {"winner", "path_len", "path" (global (row, col) pairs or None), "painted", "road"}"""
    road = _sharded_field(engine, GAME_HUNT, start_global, row0, rows_local, group)
    level = _level_of(engine, road, finish_global, start_global, row0, rows_local, group)
    if level < 0:
        engine.finish_game()
        return {"winner": -1, "path_len": 0, "path": None, "painted": 0, "road": road}
    painted = engine.paint([(level, PAINT_MASK)])
    if road == "lattice" and level > 0:
        path = _lattice_path(engine, finish_global, start_global, level, 1 << PAINT_SHIFT, want_path, row0, rows_local, group)
    else:
        path = _walk_path(engine, road, finish_global, start_global, level, 1 << PAINT_SHIFT, want_path, row0, rows_local, group)
    engine.finish_game()
    painted = sum(v[0] for v in _gather_ints(engine, [painted], group))
    return {"winner": 0, "path_len": level, "path": path, "painted": painted, "road": road}


def sharded_gather(engine, start_global, finishes_global, row0: int, rows_local: int, group=None, want_paths: bool = True):
    """Bfs::gather on this rank's band.  Collective.  -> {"claimed_by" (colour per finish), "path_len", "paths", "painted", "road"}"""
    road = _sharded_field(engine, GAME_GATHER, start_global, row0, rows_local, group)
    fins = [tuple(int(x) for x in f) for f in finishes_global]
    levels = [_level_of(engine, road, f, start_global, row0, rows_local, group) for f in fins]
    big = 1 << 40
    order = sorted(range(4), key=lambda j: (levels[j] if levels[j] >= 0 else big, fins[j][0], fins[j][1]))  # colour k claims order[k]
    rules = [(levels[order[k]] if levels[order[k]] >= 0 else INF, ((1 << k) << PAINT_SHIFT) | ((1 << k) << CACHE_SHIFT)) for k in range(4)]
    painted = engine.paint(rules)
    claimed_by = [-1] * 4
    here = lambda sq: row0 <= sq[0] < row0 + rows_local
    for k in range(4):  # the claims (bfs_threads.cc:158-161) ...
        j = order[k]
        if levels[j] >= 0:
            claimed_by[j] = k
            if here(fins[j]):
                engine.or_bits([(fins[j][0] - row0, fins[j][1])], [(1 << k) << CACHE_SHIFT])
    for k in range(4):  # ... then path.front() of every colour, in colour order (:355-364)
        j = order[k]
        if levels[j] <= 0:
            continue
        fr, fc = fins[j]
        if road == "room":  # first of N, E, S, W that is one level lower
            sr, sc = start_global
            front = (fr - 1, fc) if fr > sr else (fr, fc + 1) if fc < sc else (fr + 1, fc) if fr < sr else (fr, fc - 1)
        else:
            state = [0, 0, 0, 0]
            if here(fins[j]):
                steps, r, c, left = engine.walk(fr - row0, fc, False, True, 0, None)
                if left == levels[j]:  # no step inside the band: the front square is across the border
                    r = r - 1 if fr == row0 else r + 1
                state = [1, r + row0, c, 0]
            got = next(v for v in _gather_ints(engine, state, group) if v[0] == 1)
            front = (got[1], got[2])
        if here(front):
            engine.recolour(front[0] - row0, front[1], (1 << k) << PAINT_SHIFT)
    paths, path_len = [None] * 4, [0] * 4
    for k in range(4):
        j = order[k]
        if levels[j] >= 0:
            path_len[k] = levels[j]
            if want_paths:
                paths[k] = _walk_path(engine, road, fins[j], start_global, levels[j], 0, True, row0, rows_local, group)
    engine.finish_game()
    painted = sum(v[0] for v in _gather_ints(engine, [painted], group))
    return {"claimed_by": claimed_by, "path_len": path_len, "paths": paths, "painted": painted, "road": road}


def sharded_corners(engine, starts_global, finish_global, row0: int, rows_local: int, group=None):
    """Bfs::corners (solvers/bfs_threads.cc:420-453) on this rank's band: four colours, each from its own corner, one finish
    (the caller placed the bits and carved the finish's cross as Bfs::corners does, e.g. with lab.place_corners).  Collective.
    Four fields are searched at once by the tile wave, band by band, to their fixpoint; the canonical model then paints
    colour k on {d_k < D}, D = the lowest level at which any colour reaches the finish (winner: the lowest such colour) --
    post.cuh resolve_body's rules.  The static game never repaints; the winner's path is not returned.
    -> {"winner", "level", "painted", "rounds", "road"}"""
    starts = [tuple(int(x) for x in s_) for s_ in starts_global]
    fr, fc = (int(x) for x in finish_global)
    here = lambda r: row0 <= r < row0 + rows_local
    import torch.distributed as dist

    rank, world = dist.get_rank(group), dist.get_world_size(group)
    engine.begin_corners([(r - row0, c) if here(r) else None for r, c in starts])
    top, bottom = engine.path_rows()
    above_bottom, below_top = _exchange(dist, rank, world, top, bottom, engine, group)
    engine.set_neighbour_paths(above_bottom, below_top)
    rounds = _wave_to_fixpoint(engine, group)
    mine = [engine.level_at(fr - row0, fc, k) if here(fr) else -1 for k in range(4)]
    mine = [-1 if v == INF else v for v in mine]
    levels = [max(v[k] for v in _gather_ints(engine, mine, group)) for k in range(4)]
    reached = [v for v in levels if v >= 0]
    level = min(reached) if reached else INF
    winner = next((k for k in range(4) if levels[k] == level), -1) if reached else -1
    painted = engine.paint([(level, (1 << k) << PAINT_SHIFT, k) for k in range(4)])
    engine.finish_game()
    painted = sum(v[0] for v in _gather_ints(engine, [painted], group))
    return {"winner": winner, "level": level if reached else -1, "painted": painted, "rounds": rounds, "road": "wave"}
