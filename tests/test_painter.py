"""The heat-map painter (painters/distance.cc:45-70 `painter`) on the device: every map
square's colour from the reference's double arithmetic, bit for bit.  Checked against the oracle's
orc_painter_rgb, which tests/test_oracle_vs_reference.py pins to the colours the real reference prints.
CPU half: the kernel body on the warp emulator; GPU half: lab_distance_paint through the C ABI."""
import ctypes

import numpy as np
import pytest

from _util import Emu

INF = 0xFFFFFFFF


@pytest.fixture(scope="module")
def emu():
    return Emu()


def unpack(px):
    """pixels -> the three ints the reference prints per square, -1 where it prints a wall glyph"""
    rgb = np.stack([(px >> s) & 0xFF for s in (0, 8, 16)], axis=-1).astype(np.int32)
    rgb[(px >> 24) == 0] = -1
    assert (((px >> 24) == 0) | ((px >> 24) == 0xFF)).all()
    return rgb


def emu_pixels(emu, dist, mx, channel, nwarps=3):
    rows, cols = dist.shape
    px = np.zeros((rows, cols), np.uint32)
    d = np.ascontiguousarray(dist, np.uint32)
    assert emu.lib.emu_heatmap(rows, cols, d.ctypes.data_as(ctypes.c_void_p), ctypes.c_uint64(mx), channel, px.ctypes.data_as(ctypes.c_void_p), nwarps) == 0
    return px


@pytest.mark.parametrize("builder,rows,cols", [("kruskal", 33, 35), ("wilson", 67, 131), ("arena", 41, 57), ("rdfs", 129, 65)])
@pytest.mark.parametrize("channel", [0, 1, 2])
def test_emulated_heatmap_equals_the_reference_arithmetic(lab, O, emu, builder, rows, cols, channel):
    maze = lab.build_maze(builder, rows, cols, seed=rows)
    _, d, mx, _ = O.distance_map(maze)
    assert (unpack(emu_pixels(emu, d, mx, channel)) == O.painter_rgb(d, mx, channel)).all()


def test_emulated_heatmap_every_level_of_long_maps(O, emu):
    """All levels 0..max for maxima where 255 * (max - d) / max lands on or next to integers (the truncation
    is where a re-derived formula would differ from the reference's doubles)."""
    for mx in (1, 2, 3, 5, 17, 85, 254, 255, 256, 510, 765, 1000, 65530, 196605, 1 << 20):
        d = np.arange(0, min(mx, 70000) + 1, dtype=np.uint32)
        if mx > 70000:
            d = np.concatenate([d, np.linspace(70000, mx, 5000).astype(np.uint32), np.array([mx - 1, mx, INF], np.uint32)])
        d = np.concatenate([d, np.full((-len(d)) % 7, INF, np.uint32)]).reshape(-1, 7)
        assert (unpack(emu_pixels(emu, d, mx, 1)) == O.painter_rgb(d, mx, 1)).all(), mx


def test_emulated_heatmap_of_a_lonely_start(O, emu):
    """max == 0: the reference divides 0.0 by 0.0; its casts give dark 0, bright 128 on the x86 it runs on."""
    d = np.full((3, 5), INF, np.uint32)
    d[1, 2] = 0
    px = unpack(emu_pixels(emu, d, 0, 2))
    assert (px == O.painter_rgb(d, 0, 2)).all() and tuple(px[1, 2]) == (0, 0, 128)


# ------------------------------------------------------------------------------------------- GPU
@pytest.mark.gpu
@pytest.mark.parametrize("builder,rows,cols", [("kruskal", 1023, 1025), ("arena", 511, 2047), ("wilson", 255, 383), ("grid", 769, 513)])
def test_gpu_heatmap(lab, O, builder, rows, cols):
    maze = lab.build_maze(builder, rows, cols, seed=7)
    _, d_ref, mx_ref, _ = O.distance_map(maze)
    with lab.Grid.from_host(maze) as g:
        r = g.distance_map()
        for channel in (0, 2):
            px = g.paint(channel)
            assert (unpack(px) == O.painter_rgb(d_ref, mx_ref, channel)).all()
    assert r.max == mx_ref


@pytest.mark.gpu
def test_gpu_heatmap_needs_a_map(lab):
    maze = lab.build_maze("kruskal", 65, 65, seed=1)
    with lab.Grid.from_host(maze) as g:
        with pytest.raises(lab.LabError):
            g.paint(0)
        g.distance_map()
        with pytest.raises(lab.LabError):
            g.paint(3)
