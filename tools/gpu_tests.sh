#!/bin/bash
# GPU parity suite with hard per-test and overall timeouts (a hang must never eat the GPU budget)
export LAB_BFS_BUDGET_MS=${LAB_BFS_BUDGET_MS:-3000}
timeout ${SUITE_TIMEOUT:-420} python -m pytest tests -m gpu -x -q --timeout=${TEST_TIMEOUT:-90} "$@" 2>&1 | tail -${TAIL:-30}
