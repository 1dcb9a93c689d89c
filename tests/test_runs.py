"""Runs::paint_runs (painters/runs.cc:130-161): the straight-run length of every square's tree path.
Oracle: oracle/restate.cc orc_runs_map (FIFO search, any maze), pinned to the REAL reference text (oracle/_ref/libref_runs.so).
Product: lab_runs_map / lab_runs_paint (level field of the distance map + one pass; one tree only)."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

from _util import ROOT, Emu

INF = 0xFFFFFFFF
BIN = os.path.join(ROOT, "multithreading-with-mazes_b200", "bin")


@pytest.mark.ref
@pytest.mark.parametrize("builder,mod", [("kruskal", None), ("wilson", None), ("rdfs", None), ("arena", None), ("grid", None), ("grid", "cross"), ("kruskal", "x")])
def test_restatement_equals_the_reference(lab, ref, builder, mod):
    if not ref.have_ref_runs():
        pytest.skip("oracle/_ref/libref_runs.so not built")
    for rows, cols in [(7, 7), (31, 41), (63, 65), (101, 77)]:
        maze = lab.build_maze(builder, rows, cols, seed=rows * cols, mod=mod)
        g1, r1, mx1 = ref.ref_runs(maze, seed=3)
        g2, r2, mx2, settled = ref.runs_map(maze)
        r1 = np.where(r1 == np.uint64(2**64 - 1), INF, r1).astype(np.uint32)
        assert (r1 == r2).all() and (g1 == g2).all() and mx1 == mx2
        assert settled == int((r2 != INF).sum())


@pytest.mark.parametrize("builder,rows,cols", [("kruskal", 63, 65), ("wilson", 129, 67), ("rdfs", 65, 131), ("kruskal", 9, 9)])
def test_emulated_runs_from_the_level_field(lab, O, builder, rows, cols):
    """runs_body on the warp emulator: on one tree the run lengths follow from the oracle's level field."""
    maze = lab.build_maze(builder, rows, cols, seed=11)
    _, d_ref, _, _ = O.distance_map(maze)
    _, r_ref, mx_ref, _ = O.runs_map(maze)
    emu = Emu()
    runs = np.zeros((rows, cols), np.uint32)
    mx = ctypes.c_uint32()
    emu.lib.emu_runs(rows, cols, d_ref.ctypes.data_as(ctypes.c_void_p), runs.ctypes.data_as(ctypes.c_void_p), ctypes.byref(mx), 3)
    assert (runs == r_ref).all() and mx.value == mx_ref


@pytest.mark.gpu
@pytest.mark.parametrize("builder,n", [("kruskal", 255), ("wilson", 1023), ("rdfs", 767), ("kruskal", 2047)])
def test_gpu_runs_map_and_painter(lab, O, builder, n):
    maze = lab.build_maze(builder, n, n + 2, seed=0xB2000002)
    with lab.Grid.from_host(maze) as g:
        r = g.runs_map()
        after = g.download()
        pixels = g.paint_runs(2)
    g_ref, r_ref, mx_ref, settled_ref = O.runs_map(maze)
    assert (r.dist == r_ref).all() and r.max == mx_ref and r.settled == settled_ref and (after == g_ref).all()
    rgb = np.stack([(pixels >> sh) & 0xFF for sh in (0, 8, 16)], axis=-1).astype(np.int32)
    rgb[(pixels >> 24) == 0] = -1
    assert (rgb == O.painter_rgb(r_ref, mx_ref, 2)).all()  # (runs.cc:45-70 is distance.cc:45-70's arithmetic)


@pytest.mark.gpu
def test_gpu_runs_needs_one_tree(lab):
    for maze in (lab.build_maze("arena", 65, 65, seed=1), lab.build_maze("grid", 65, 67, seed=1, mod="cross")):
        with lab.Grid.from_host(maze) as g:
            with pytest.raises(lab.LabError) as e:
                g.runs_map()
            assert e.value.code == -1 and "one tree" in str(e.value)


@pytest.mark.gpu
def test_cli_measure_runs_and_animated_variants(lab, O, tmp_path):
    """measure -p runs leaves the reference's grid; the animated variants (-pa / -sa: frames replayed from the GPU's level
    fields) leave the same grids as the static ones and draw one cursor-addressed square per painted square."""
    seed = 4242
    out = str(tmp_path / "o.labmaze")
    env = dict(os.environ, LAB_SEED=str(seed), LAB_DUMP=out, LAB_QUIET="1")
    r = subprocess.run([os.path.join(BIN, "measure"), "-r", "63", "-c", "129", "-b", "kruskal", "-p", "runs"], env=env, capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr
    maze = lab.build_maze("kruskal", 63, 129, seed=seed)
    assert (lab.read_maze(out) == O.runs_map(maze)[0]).all()
    env.pop("LAB_QUIET")
    for tool, args in [("measure", ["-p", "distance", "-pa", "7"]), ("measure", ["-p", "runs", "-pa", "7"]), ("run_maze", ["-s", "bfs-hunt", "-sa", "7"]),
                       ("run_maze", ["-s", "bfs-gather", "-sa", "7"]), ("run_maze", ["-s", "bfs-corners", "-sa", "7"]),
                       ("run_maze", ["-s", "darkbfs-hunt", "-sa", "7"])]:
        static_args = [a for a in args if a not in ("-pa", "-sa", "7")]
        a = subprocess.run([os.path.join(BIN, tool), "-r", "31", "-c", "41", "-b", "wilson"] + args, env=env, capture_output=True, text=True, timeout=120)
        assert a.returncode == 0, a.stderr
        animated = lab.read_maze(out)
        b = subprocess.run([os.path.join(BIN, tool), "-r", "31", "-c", "41", "-b", "wilson"] + static_args, env=env, capture_output=True, text=True, timeout=120)
        assert b.returncode == 0, b.stderr
        static = lab.read_maze(out)
        if "bfs-corners" in args:  # (the animated corners carves 7 neighbours of the centre, the static one 4: bfs_threads.cc:428-431 vs :469-472)
            assert ((animated & 0xC000) == (static & 0xC000)).all()
        else:
            assert (animated == static).all(), (tool, args)
        painted = int(((animated & 0x0FF0) != 0).sum()) if tool == "run_maze" else int(((animated & 0x0200) != 0).sum())
        assert a.stdout.count("\033[") > painted  # every painted square was drawn at its own cursor position
