"""The sharded paths over NCCL, one rank per GPU (needs >= 2 GPUs on the box; skipped otherwise): tools/run_bands.py --check
verifies every band of a sharded distance map, hunt, gather and corners against the oracle's canonical model on the whole maze.
(The same commands were run by hand on 2 and 8 B200s: profiles/r02b_bands_general_and_corners_n2.md, r02b_hunt_lattice_bands_n2.log.)"""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CASES = [
    ["--builder", "kruskal", "--rows", "2047", "--cols", "1023", "--super"],                       # distance map, lattice road across bands
    ["--builder", "wilson", "--rows", "1023", "--cols", "515"],                                    # distance map, general pipeline across bands
    ["--builder", "grid", "--mod", "cross", "--rows", "1023", "--cols", "1023"],                    # distance map, the wave band by band
    ["--game", "hunt", "--builder", "kruskal", "--rows", "2047", "--cols", "1023", "--super"],      # hunt, path without hand-overs
    ["--game", "gather", "--builder", "kruskal", "--mod", "x", "--rows", "1023", "--cols", "515"],  # gather on a maze with loops
    ["--game", "corners", "--builder", "grid", "--mod", "cross", "--rows", "1023", "--cols", "1023"],
]


def _gpus():
    try:
        import torch

        return torch.cuda.device_count()
    except Exception:
        return 0


@pytest.mark.gpu
@pytest.mark.parametrize("case", CASES, ids=lambda c: "-".join(x.strip("-") for x in c))
def test_sharded_over_nccl(case):
    n = _gpus()
    if n < 2:
        pytest.skip("needs at least two GPUs")
    world = 2
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1", "--master-port",
           str(29700 + (os.getpid() % 200)), os.path.join(ROOT, "tools", "run_bands.py"), "--check", "--reps", "1"] + case
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    text = out.stdout + out.stderr
    assert out.returncode == 0, text[-2000:]
    oks = [l for l in text.splitlines() if l.startswith("rank ") and l.rstrip().endswith(" OK")]
    assert len(oks) == world and "MISMATCH" not in text, text[-2000:]
