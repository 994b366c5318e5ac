#!/bin/bash
# round 2, session B: device-side task glue on the GPU (no host synchronisation allowed) + smoke
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
L=gpurun_out/r2b_glue.log
: > $L
timeout 300 python -m pytest tests/test_task_glue.py -q 2>&1 | tail -3 >> $L
timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2 >> $L
cat $L
