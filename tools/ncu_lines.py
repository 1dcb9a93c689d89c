"""Developer tool: per-source-line instruction counts / stall samples of one kernel from an .ncu-rep.

ncu's CSV export of the source page is SASS-only; this joins it with `nvdisasm --print-line-info`
of the cubin inside liblabyrinth_b200.so (same build!) through the instruction offsets.
usage: python tools/ncu_lines.py REPORT.ncu-rep KERNEL_REGEX [top_n]"""
import csv, io, os, re, subprocess, sys, tempfile, collections

rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "multithreading-with-mazes_b200", "liblabyrinth_b200.so")
tmp = tempfile.mkdtemp()
subprocess.check_call(["cuobjdump", "-xelf", "all", so], cwd=tmp, stdout=subprocess.DEVNULL)
cubin = [f for f in os.listdir(tmp) if f.startswith("lab.")][0]
dis = subprocess.run(["nvdisasm", "--print-line-info", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout
# offset -> (file, line) for the wanted function
line_of, cur, infn = {}, None, False
for l in dis.split("\n"):
    m = re.match(r"\s*\.section\s+\.text\.(\S+?),", l)
    if m:
        infn = re.search(kern, m.group(1)) is not None
        continue
    if not infn:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/", l)
    if m and cur:
        line_of[int(m.group(1), 16)] = cur
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kern], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = next(r for r in rows if r and r[0] == "Address")
ia, ii, isamp = hdr.index("Address"), hdr.index("Instructions Executed"), hdr.index("# Samples")
ith = hdr.index("Thread Instructions Executed")
agg = collections.defaultdict(lambda: [0, 0, 0])
base = None
for r in rows:
    if len(r) != len(hdr) or not r[ia].startswith("0x"):
        continue
    a = int(r[ia], 16)
    base = a if base is None else base
    k = line_of.get(a - base, ("?", 0))
    agg[k][0] += int(r[ii]); agg[k][1] += int(r[isamp]); agg[k][2] += int(r[ith])
tot = sum(v[0] for v in agg.values()); st = sum(v[1] for v in agg.values())
print(f"total warp instructions {tot}, samples {st}")
src = {}
for (f, n), v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    if f not in src:
        for d in ("multithreading-with-mazes_b200/csrc/device", "multithreading-with-mazes_b200/csrc"):
            p = os.path.join(ROOT, d, f)
            if os.path.exists(p):
                src[f] = open(p).read().split("\n"); break
        else:
            src[f] = []
    text = src[f][n - 1].strip()[:80] if 0 < n <= len(src[f]) else ""
    print(f"{f}:{n:<5d} inst {100*v[0]/tot:5.1f}%  stall {100*v[1]/max(st,1):5.1f}%  thr/inst {v[2]/max(v[0],1):4.1f}  {text}")
