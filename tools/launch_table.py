"""Developer tool: ncu launch-list CSV (gpu__time_duration.sum, dram__bytes_read.sum, dram__bytes_write.sum) -> markdown table.
usage: python tools/launch_table.py LAUNCHES.csv "title" "command" [alg_bytes] > profiles/xxx.md"""
import collections, csv, sys

path, title, command = sys.argv[1], sys.argv[2], sys.argv[3]
alg = float(sys.argv[4]) if len(sys.argv) > 4 else None
rows = [r for r in csv.reader(l for l in open(path) if l.startswith('"'))]
hdr = rows[0]
ik, im, iv, iid = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("ID")
per = collections.OrderedDict()
for r in rows[1:]:
    d = per.setdefault(r[iid], {"k": r[ik]})
    d[r[im]] = float(r[iv].replace(",", ""))
agg = collections.OrderedDict()
for d in per.values():
    name = d["k"]
    lib = name.startswith("k_") or "k_tree" in name or "lab::" in name or "<unnamed>::k_" in name
    short = name.split("(")[0].replace("void ", "").replace("(anonymous namespace)::", "").replace("<unnamed>::", "")
    if not lib:
        short = "(caller) " + short[:60]
    a = agg.setdefault(short, {"n": 0, "us": 0.0, "rd": 0.0, "wr": 0.0, "lib": lib})
    a["n"] += 1
    a["us"] += d.get("gpu__time_duration.sum", 0.0) / 1e3
    a["rd"] += d.get("dram__bytes_read.sum", 0.0) / 1e6
    a["wr"] += d.get("dram__bytes_write.sum", 0.0) / 1e6
lib_us = sum(a["us"] for a in agg.values() if a["lib"])
print(f"# {title}\n")
print(f"Command: `{command}`\n(after the same command exited 0 without ncu).  Times are cold-cache and serialised: compare SHARES, not absolutes.  DRAM bytes are per-launch averages.\n")
print("| kernel | launches | total us | share of library kernels | avg us | avg DRAM read MB | avg DRAM write MB |")
print("|---|---:|---:|---:|---:|---:|---:|")
for k, a in agg.items():
    share = f"{100 * a['us'] / lib_us:.1f}%" if a["lib"] and lib_us else "-"
    print(f"| `{k}` | {a['n']} | {a['us']:.1f} | {share} | {a['us'] / a['n']:.1f} | {a['rd'] / a['n']:.2f} | {a['wr'] / a['n']:.2f} |")
n_search = max((a["n"] for k, a in agg.items() if k.startswith("k_pack")), default=1)
tot = sum(a["rd"] + a["wr"] for a in agg.values() if a["lib"]) / n_search
print(f"\nLibrary kernels: {lib_us / n_search:.1f} us and {tot:.1f} MB of DRAM traffic per search ({n_search} searches captured)" +
      (f", against {alg / 1e6:.1f} MB algorithmic." if alg else "."))
