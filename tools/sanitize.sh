#!/bin/bash
# compute-sanitizer memcheck + racecheck over one small search per road (lattice, general tree pipeline, open room,
# tile wave) and one gather / hunt.  Output: gpurun_out/<tag>_sanitize_*.log (summaries are copied to profiles/).
tag=${1:-r02}
mkdir -p gpurun_out
run() { # name, env, args...
  name=$1; shift; envs=$1; shift
  for tool in memcheck racecheck; do
    env $envs timeout 600 compute-sanitizer --tool $tool --print-limit 5 python tools/run_one.py "$@" > gpurun_out/${tag}_sanitize_${name}_${tool}.log 2>&1
    echo "$name $tool: $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' gpurun_out/${tag}_sanitize_${name}_${tool}.log | tail -1)"
  done
}
run lattice_distance "LAB_X=1" distance kruskal 383 1
run lattice_gather "LAB_X=1" gather wilson 383 1
run tree_distance "LAB_LATTICE=0" distance kruskal 383 1
run tree_hunt "LAB_LATTICE=0" hunt kruskal 383 1
run room_gather "LAB_X=1" gather arena 383 1
run wave_distance "LAB_X=1" distance grid 383 1
