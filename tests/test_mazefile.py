"""The raw maze file (include/labyrinth_b200.h lab_maze_*): what makes "identical serialized mazes" a file."""
import os
import subprocess

import numpy as np
import pytest

from _util import ROOT

BIN = os.path.join(ROOT, "multithreading-with-mazes_b200", "bin")


def test_round_trip_and_layout(lab, tmp_path):
    maze = lab.build_maze("kruskal", 33, 47, seed=5)
    path = str(tmp_path / "m.labmaze")
    lab.write_maze(path, maze)
    raw = open(path, "rb").read()
    assert raw[:8] == b"LABMAZE1" and len(raw) == 16 + maze.size * 2
    assert tuple(np.frombuffer(raw[8:16], "<i4")) == (33, 47)
    assert (np.frombuffer(raw[16:], "<u2").reshape(33, 47) == maze).all()  # dense row-major, little endian
    assert lab.maze_header(path) == (33, 47)
    assert (lab.read_maze(path) == maze).all()
    assert (np.asarray(lab.read_maze(path, mmap=True)) == maze).all()
    assert not os.path.exists(path + ".tmp")


def test_errors(lab, tmp_path):
    with pytest.raises(lab.LabError) as e:
        lab.maze_header(str(tmp_path / "nope"))
    assert e.value.code == -6
    bad = tmp_path / "bad"
    bad.write_bytes(b"NOTAMAZE" + b"\0" * 32)
    with pytest.raises(lab.LabError):
        lab.read_maze(str(bad))
    maze = lab.build_maze("arena", 9, 9, seed=1)
    path = str(tmp_path / "m")
    lab.write_maze(path, maze)
    with open(path, "r+b") as f:  # truncated
        f.truncate(16 + 50)
    with pytest.raises(lab.LabError):
        lab.read_maze(path)
    with pytest.raises(lab.LabError):
        lab.write_maze(str(tmp_path / "no" / "such" / "dir" / "m"), maze)


def test_cache(lab, tmp_path):
    d = str(tmp_path / "cache")
    a = lab.cached_maze("wilson", 31, 33, seed=9, cache_dir=d)
    files = os.listdir(d)
    assert len(files) == 1 and files[0].startswith("wilson-none-31x33-")
    b = lab.cached_maze("wilson", 31, 33, seed=9, cache_dir=d)
    assert (a == b).all() and (a == lab.build_maze("wilson", 31, 33, seed=9)).all()
    c = lab.cached_maze("wilson", 31, 33, seed=10, cache_dir=d)
    assert len(os.listdir(d)) == 2 and (c != a).any()


def test_oracle_reads_the_same_bytes(lab, O, tmp_path):
    """The oracle consumes the file the product wrote (np.fromfile at the documented offset)."""
    maze = lab.build_maze("kruskal", 63, 65, seed=3)
    path = str(tmp_path / "m.labmaze")
    lab.write_maze(path, maze)
    got = O.read_maze_file(path)
    assert (got == maze).all()
    assert O.distance_map(got)[3] == O.distance_map(maze)[3]


@pytest.mark.gpu
def test_cli_load_solves_the_file(lab, O, tmp_path):
    """LAB_LOAD: run_maze / measure solve a serialized maze as it is; LAB_DUMP writes the same format back."""
    maze = lab.build_maze("wilson", 63, 129, seed=21)
    src, out = str(tmp_path / "in.labmaze"), str(tmp_path / "out.labmaze")
    lab.write_maze(src, maze)
    env = dict(os.environ, LAB_LOAD=src, LAB_DUMP=out, LAB_QUIET="1")
    r = subprocess.run([os.path.join(BIN, "measure"), "-p", "distance"], env=env, capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr
    got = lab.read_maze(out)
    assert (got == O.distance_map(maze)[0]).all()
    seed = 77
    env = dict(os.environ, LAB_LOAD=src, LAB_DUMP=out, LAB_QUIET="1", LAB_SEED=str(seed))
    r = subprocess.run([os.path.join(BIN, "run_maze"), "-s", "bfs-gather"], env=env, capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr
    placed, start, finishes = lab.place_gather(maze, seed=seed)
    assert (lab.read_maze(out) == O.canon_gather(placed, start, finishes)[0]).all()
