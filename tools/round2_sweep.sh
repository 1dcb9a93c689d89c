python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | tee gpurun_out/r03d_gputests.log
for w in c2-distance-kruskal-4095 gather-kruskal-4095 hunt-wilson-4095 distance-wilson-4095 c4-corners-grid-cross-4095 c3-gather-arena-16383 distance-arena-16383 distance-kruskal-16383; do
  timeout 300 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r03d_bench_$w.json 2> gpurun_out/r03d_bench_$w.err || echo "bench $w failed"
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r03d_bench_$w.json").read().strip().split("\n")[-1])
    print("$w", "ms/step %.3f"%d["ms_per_step"], "value %.2f"%d["value"], "frac %.4f"%d["roofline"]["frac"], "whole %.4f"%d["roofline"]["whole_step"]["frac"], "e2e %.2f"%d["e2e"]["value"], "launches", d["gpu_launches"], "road", d["search_stats_last_step"]["tree_used"])
except Exception as e:
    print("$w unreadable", e)
PY
done
