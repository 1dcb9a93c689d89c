#!/bin/bash
# usage: tools/bench_set.sh TAG workload...   -> gpurun_out/TAG_bench_<workload>.json (+ .err)
tag=$1; shift
for w in "$@"; do
  python bench.py --workload "$w" --no-cpu-baseline > "gpurun_out/${tag}_bench_${w}.json" 2> "gpurun_out/${tag}_bench_${w}.err" || echo "bench $w failed"
  python - "$w" "gpurun_out/${tag}_bench_${w}.json" <<'PY'
import json, sys
try:
    j = json.load(open(sys.argv[2]))
    print(sys.argv[1], "ms/step %.3f" % j["ms_per_step"], "value %.1f" % j["value"], "frac %.3f" % j["roofline"]["frac"],
          {k: round(v, 3) for k, v in j["phases_ms_per_step"].items()}, "used", j["search_stats_last_step"]["tree_used"])
except Exception as e:
    print(sys.argv[1], "no line:", e)
PY
done
