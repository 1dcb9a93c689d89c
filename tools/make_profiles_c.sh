#!/bin/bash
# Round artefacts (second set, tag r01c): bench lines of every workload after the open-room path, the 16-byte
# pack and the sharded tree pipeline; ncu launch list of the default bench command; full captures of the new
# streaming kernels.  Output: gpurun_out/<tag>_*.
tag=${1:-r01c}
mkdir -p gpurun_out
run() {
  timeout 900 python bench.py --steps ${STEPS:-10} --warmup 3 --workload "$1" $2 > gpurun_out/${tag}_bench_$1.json 2> gpurun_out/${tag}_bench_$1.err || echo "bench $1 failed"
  python - "$1" "gpurun_out/${tag}_bench_$1.json" <<'PY'
import json, sys
try:
    j = json.loads([l for l in open(sys.argv[2]) if l.startswith("{")][-1])
    print(sys.argv[1], "ms/step %.3f" % j["ms_per_step"], "value %.1f" % j["value"], "e2e %.2f" % j["e2e"]["value"], "frac %.3f" % j["roofline"]["frac"],
          {k: round(v, 3) for k, v in j["phases_ms_per_step"].items()}, "used", j["search_stats_last_step"]["tree_used"])
except Exception as e:
    print(sys.argv[1], "no line:", e)
PY
}
run c2-distance-kruskal-4095 ""
run distance-wilson-4095 ""
run gather-kruskal-4095 ""
run c4-corners-grid-cross-4095 ""
run distance-arena-16383 "--no-cpu-baseline"
run c3-gather-arena-16383 "--no-cpu-baseline"
run h-distance-arena-32767 "--no-cpu-baseline"
run h-gather-arena-32767 "--no-cpu-baseline"
run h-distance-arena-pillar-32767 "--no-cpu-baseline"
run h-gather-arena-pillar-32767 "--no-cpu-baseline"
STEPS=5 run distance-kruskal-16383 "--no-cpu-baseline"
if [ -z "$QUICK" ]; then STEPS=3 run h-distance-kruskal-32767 "--no-cpu-baseline"; fi
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${tag}_bench_reference.json 2> gpurun_out/${tag}_bench_reference.err
for c in "kruskal 4095" "arena 16383"; do python tools/time_paint.py $c; done > gpurun_out/${tag}_paint.jsonl 2>&1
# launch list of the default bench command (after it exited 0 without ncu above)
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 600 --csv \
  --log-file gpurun_out/${tag}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_ncu_launches.log 2>&1
# launch list of one open-room distance map and one open-room gather at the headline size
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 200 --csv \
  --log-file gpurun_out/${tag}_launches_room_distance.csv python tools/run_one.py distance arena 32767 2 > gpurun_out/${tag}_ncu_room_d.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 200 --csv \
  --log-file gpurun_out/${tag}_launches_room_gather.csv python tools/run_one.py gather arena 32767 2 > gpurun_out/${tag}_ncu_room_g.log 2>&1
# full captures of the streaming kernels the open-room path consists of
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_room_distance|k_pack|k_cross|k_tree_count" -c 4 \
  -o gpurun_out/${tag}_room_distance python tools/run_one.py distance arena 32767 1 > gpurun_out/${tag}_ncu_full_d.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_paint" -c 1 \
  -o gpurun_out/${tag}_room_gather python tools/run_one.py gather arena 32767 1 > gpurun_out/${tag}_ncu_full_g.log 2>&1
ls -la gpurun_out | tail -8
