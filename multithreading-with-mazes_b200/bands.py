"""Row-band sharding of the distance map across GPUs.

One process per GPU (``torch.distributed``, NCCL over NVLink/NVSwitch); rank r holds a contiguous
band of rows of the maze.  The search itself is local (the persistent tile kernel runs each band to
its local fixpoint); between two local searches neighbouring ranks exchange ONE ROW of levels
each way (cols x 4 bytes) and all ranks agree, with a one-word all-reduce, whether any border tile
was re-activated.  The loop ends when nobody was.  Levels are global BFS levels, so the result is
bit-identical to the single-GPU search.

``torch`` is plumbing here (device buffers for the exchanged rows, the process group); every
kernel is the library's own.
"""
from __future__ import annotations

import ctypes
from ctypes import byref, c_uint32, c_uint64, c_void_p
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


class GpuBandEngine:
    """One band on one GPU: ctypes over the lab_band_* entry points, row buffers are torch tensors."""

    def __init__(self, lab, squares, device=None):
        import torch

        self.lab, self.torch = lab, torch
        self.squares = squares  # int16 CUDA tensor [rows_local, cols], owned by the caller
        rows, cols = squares.shape
        self.grid = lab.Grid.from_device_ptr(rows, cols, squares.data_ptr(), keepalive=squares)
        self.grid.set_stream(torch.cuda.current_stream().cuda_stream)
        self.L = lab.lib()
        self.h = self.grid._h
        self.cols_padded = int(self.L.lab_band_cols_padded(self.h))
        dev = squares.device
        self._u8 = [torch.empty(self.cols_padded, dtype=torch.uint8, device=dev) for _ in range(2)]
        self._u32 = [torch.empty(self.cols_padded, dtype=torch.int32, device=dev) for _ in range(2)]
        self.row_device = dev

    def _check(self, rc):
        if rc != 0:
            raise self.lab.LabError(rc, self.L.lab_last_error().decode(errors="replace"))

    @staticmethod
    def _ptr(t):
        return c_void_p(t.data_ptr()) if t is not None else None

    def begin(self, src_local: Optional[Tuple[int, int]]):
        r, c = src_local if src_local is not None else (-1, -1)
        self._check(self.L.lab_band_begin(self.h, r, c))

    def path_rows(self):
        self._check(self.L.lab_band_path_rows(self.h, self._ptr(self._u8[0]), self._ptr(self._u8[1])))
        return self._u8[0], self._u8[1]

    def set_neighbour_paths(self, above_bottom, below_top):
        self._check(self.L.lab_band_set_neighbour_paths(self.h, self._ptr(above_bottom), self._ptr(below_top)))

    def relax(self):
        self._check(self.L.lab_band_relax(self.h))

    def export_edges(self):
        self._check(self.L.lab_band_export_edges(self.h, self._ptr(self._u32[0]), self._ptr(self._u32[1])))
        return self._u32[0], self._u32[1]

    def import_edges(self, above_bottom, below_top) -> int:
        n = c_uint32()
        self._check(self.L.lab_band_import_edges(self.h, self._ptr(above_bottom), self._ptr(below_top), byref(n)))
        return int(n.value)

    def finish(self) -> Tuple[int, int]:
        mx, st = c_uint64(), c_uint64()
        self._check(self.L.lab_band_finish(self.h, byref(mx), byref(st)))
        return int(mx.value), int(st.value)

    def new_row(self, like):
        return self.torch.empty_like(like)


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


def sharded_distance_map(engine, start_global: Tuple[int, int], row0: int, rows_local: int, group=None, max_rounds: int = 1 << 20):
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
    while True:
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
    mx_t = agg[:1].clone()
    st_t = agg[1:].clone()
    dist.all_reduce(mx_t, op=dist.ReduceOp.MAX, group=group)
    dist.all_reduce(st_t, op=dist.ReduceOp.SUM, group=group)
    return {"max": int(mx_t.item()), "settled": int(st_t.item()), "rounds": rounds}
