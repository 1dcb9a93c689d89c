#!/bin/bash
# usage: tools/scale_run.sh TAG N...   (inside a gpurun --gpus MAX call) -> gpurun_out/TAG_scale_N.json
tag=$1; shift
for n in "$@"; do
  if [ "$n" = 1 ]; then
    python bench.py --gpus 1 --no-cpu-baseline > gpurun_out/${tag}_scale_1.json 2> gpurun_out/${tag}_scale_1.err
  else
    python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29700 + n)) bench.py --gpus $n --no-cpu-baseline > gpurun_out/${tag}_scale_$n.json 2> gpurun_out/${tag}_scale_$n.err
  fi
  python - gpurun_out/${tag}_scale_$n.json <<'PY'
import json, sys
try:
    j = json.loads([l for l in open(sys.argv[1]) if l.startswith("{")][-1])
    print(j["n_gpus"], "GPUs: ms/step %.3f" % j["ms_per_step"], "value %.2f Gcells/s" % j["value"], "kernel_ms %.3f" % j["roofline"]["kernel_ms"], "e2e %.2f" % j["e2e"]["value"], j["config"]["rows"], "x", j["config"]["cols"])
except Exception as e:
    print(sys.argv[1], "no line:", e)
PY
done
