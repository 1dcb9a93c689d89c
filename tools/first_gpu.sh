#!/bin/bash
# first GPU contact: parity tests + a timing table
set -x
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -25
timeout 600 python tools/timing_table.py 2>&1 | tee gpurun_out/timing_table.txt
