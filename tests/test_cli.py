"""The reference-facing C++ layer: include/labyrinth.hh (Bfs::*, Distance::* with the reference's
signatures) driven through the two CLIs with the reference's flags."""
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "multithreading-with-mazes_b200", "bin")


def run(prog, *args, env=None, check=False):
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run([os.path.join(BIN, prog), *args], capture_output=True, text=True, env=e, check=check, timeout=120)


@pytest.mark.parametrize("prog", ["run_maze", "measure"])
def test_help_and_flag_errors(prog):
    assert run(prog, "-h").returncode == 0
    bad = run(prog, "-z")
    assert bad.returncode == 1 and "Invalid argument flag: -z" in bad.stderr
    small = run(prog, "-r", "5")  # < 7 is rejected (run_maze.cc:280-299)
    assert small.returncode == 1 and "Invalid argument: 5" in small.stderr
    unknown = run(prog, "-b", "nonesuch")
    assert unknown.returncode == 1


def test_out_of_scope_solvers_are_argument_errors():
    assert run("run_maze", "-s", "dfs-hunt").returncode == 1
    assert run("measure", "-p", "runs2").returncode == 1


def read_dump(path):
    raw = open(path, "rb").read()
    assert raw[:8] == b"LABMAZE1"
    rows, cols = np.frombuffer(raw[8:16], np.int32)
    return np.frombuffer(raw[16:], np.uint16).reshape(rows, cols).copy()


@pytest.mark.gpu
def test_run_maze_config1_matches_oracle(lab, O, tmp_path):
    """BASELINE config 1: run_maze -r 51 -c 111 -b rdfs -s bfs-hunt."""
    dump = str(tmp_path / "hunt.bin")
    seed = 0xB2000001
    out = run("run_maze", "-r", "51", "-c", "111", "-b", "rdfs", "-s", "bfs-hunt", env={"LAB_SEED": str(seed), "LAB_DUMP": dump})
    assert out.returncode == 0, out.stderr
    assert "thread won!" in out.stdout
    got = read_dump(dump)
    maze = lab.build_maze("rdfs", 51, 111, seed)          # the builder drew device seed `seed`
    placed, s, f = lab.place_hunt(maze, seed + 1)          # the solver the next two
    ref, winner, path = O.canon_hunt(placed, s, f)
    assert (got == ref).all() and winner == 0


@pytest.mark.gpu
@pytest.mark.parametrize("solver,game", [("bfs-gather", "gather"), ("bfs-corners", "corners"), ("darkbfs-hunt", "hunt")])
def test_run_maze_other_games(lab, O, tmp_path, solver, game):
    dump = str(tmp_path / "g.bin")
    seed = 77
    out = run("run_maze", "-r", "65", "-c", "131", "-b", "kruskal", "-m", "cross", "-s", solver, env={"LAB_SEED": str(seed), "LAB_DUMP": dump, "LAB_QUIET": "1"})
    assert out.returncode == 0, out.stderr
    got = read_dump(dump)
    maze = lab.build_maze("kruskal", 65, 131, seed, mod="cross")
    if game == "gather":
        placed, s, fin = lab.place_gather(maze, seed + 1)
        ref = O.canon_gather(placed, s, fin)[0]
    elif game == "corners":
        placed, st, f = lab.place_corners(maze, seed + 1)
        ref = O.canon_corners(placed, st, f)[0]
    else:
        placed, s, f = lab.place_hunt(maze, seed + 1)
        ref = O.canon_hunt(placed, s, f)[0]
    assert (got == ref).all()


@pytest.mark.gpu
def test_measure_distance_matches_oracle(lab, O, tmp_path):
    dump = str(tmp_path / "d.bin")
    seed = 5
    out = run("measure", "-r", "63", "-c", "129", "-b", "wilson", "-p", "distance", env={"LAB_SEED": str(seed), "LAB_DUMP": dump})
    assert out.returncode == 0, out.stderr
    got = read_dump(dump)
    maze = lab.build_maze("wilson", 63, 129, seed)
    g_ref, d_ref, mx, settled = O.distance_map(maze)
    assert (got == g_ref).all()
    # the painted colours are the reference painter's arithmetic on the GPU's levels
    import re
    cells = re.findall(r"\x1b\[(\d+);(\d+)f\x1b\[38;2;(\d+);(\d+);(\d+)m", out.stdout)
    rgb = {(int(r) - 1, int(c) - 1): (int(a), int(b), int(cc)) for r, c, a, b, cc in cells}
    assert len(rgb) == settled
    expect = [O.painter_rgb(d_ref, mx, ch) for ch in range(3)]
    assert any(all(tuple(e[r, c]) == v for (r, c), v in rgb.items()) for e in expect)
