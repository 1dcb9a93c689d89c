"""The C-ABI library: loads, exports every symbol the header declares, and fails loudly
(no CPU fallback) when asked to compute without a GPU."""
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    text = open(os.path.join(ROOT, "include", "labyrinth_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(lab_[a-z_0-9]+)\s*\(", text)))


def test_every_declared_symbol_is_exported(lab):
    L = lab.lib()
    syms = header_symbols()
    assert len(syms) >= 25
    missing = [s for s in syms if not hasattr(L, s)]
    assert not missing, missing
    assert set(lab.ABI_SYMBOLS) == set(syms)


def test_library_is_built_for_sm_100a(lab):
    import subprocess

    out = subprocess.run(["cuobjdump", "-lelf", lab.LIB_PATH], capture_output=True, text=True)
    if out.returncode != 0:
        pytest.skip("cuobjdump not available")
    assert "sm_100a" in out.stdout


def test_no_cpu_fallback(lab):
    n, info = lab.device_info()
    if n > 0:
        pytest.skip("a GPU is present")
    assert "no CPU fallback" in info
    with pytest.raises(lab.LabError) as e:
        lab.Grid.from_host(np.zeros((9, 9), np.uint16))
    assert e.value.code == -2  # LAB_ERR_CUDA


def test_host_helpers_validate_arguments(lab):
    with pytest.raises(lab.LabError):
        lab.build_maze("kruskal", 8, 9, 1)  # even size
    with pytest.raises(lab.LabError):
        lab.build_maze("nonesuch", 9, 9, 1)
    with pytest.raises(lab.LabError):
        lab.build_maze("arena", 9, 9, 1, mod="diagonal")
    assert lab.distance_start(4095, 4095) == (2047, 2047)
    assert lab.distance_start(65535, 65535) == (32767, 32767)
    assert lab.distance_start(9, 9) == (5, 5)


def test_product_does_not_touch_the_oracle():
    """Nothing under the package may import, link or open anything under oracle/."""
    pkg = os.path.join(ROOT, "multithreading-with-mazes_b200")
    for dirpath, _, files in os.walk(pkg):
        if os.sep + "build" in dirpath:
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cc", ".hh", ".h", "Makefile")):
                text = open(os.path.join(dirpath, f), errors="replace").read()
                for pat in (r'#\s*include\s*[<"][^>"]*oracle', r"^\s*(from|import)\s+oracle", r"liboracle", r"libref_", r"CDLL\([^)]*oracle"):
                    assert not re.search(pat, text, flags=re.M), (os.path.join(dirpath, f), pat)
