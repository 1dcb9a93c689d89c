#!/bin/bash
# Round artefacts on the GPU box: bench lines of every workload, the ncu launch list of the default
# bench command and one full capture of the pipeline's kernels.  Output: gpurun_out/<tag>_*.
tag=${1:-r01}
mkdir -p gpurun_out
run() { # name, extra args
  timeout 900 python bench.py --steps ${STEPS:-10} --warmup 3 --workload "$1" $2 > gpurun_out/${tag}_bench_$1.json 2> gpurun_out/${tag}_bench_$1.err || echo "bench $1 failed"
  tail -c 600 gpurun_out/${tag}_bench_$1.json | head -c 0; python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/${tag}_bench_$1.json").read().strip().split("\n")[-1])
    print("$1", "value %.2f Gc/s"%d["value"], "e2e %.2f"%d["e2e"]["value"], "ms/step %.3f"%d["ms_per_step"], "frac %.4f"%d["roofline"]["frac"], "tree", d["search_stats_last_step"]["tree_used"])
except Exception as e:
    print("$1 unreadable", e)
PY
}
run c2-distance-kruskal-4095 ""
run distance-wilson-4095 ""
run gather-kruskal-4095 ""
run c4-corners-grid-cross-4095 ""
run distance-arena-16383 "--no-cpu-baseline"
run c3-gather-arena-16383 "--no-cpu-baseline"
if [ -z "$QUICK" ]; then
  run h-distance-arena-32767 "--no-cpu-baseline"
  run h-gather-arena-32767 "--no-cpu-baseline"
  STEPS=5 run distance-kruskal-16383 "--no-cpu-baseline"
  STEPS=3 run h-distance-kruskal-32767 "--no-cpu-baseline"
fi
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${tag}_bench_reference.json 2> gpurun_out/${tag}_bench_reference.err
# launch list of the default bench command (after it exited 0 without ncu above)
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 600 --csv \
  --log-file gpurun_out/${tag}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_ncu_launches.log 2>&1
# one full capture of the pipeline's kernels (first search of the run)
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_tree_(walk|flood|jump|fixup)" -c 5 \
  -o gpurun_out/${tag}_tree_kernels python tools/run_one.py distance kruskal 4095 1 > gpurun_out/${tag}_ncu_full.log 2>&1
ls -la gpurun_out | tail -5
