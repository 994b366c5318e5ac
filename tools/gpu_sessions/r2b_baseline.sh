#!/bin/bash
# round 2, session B: state of HEAD on hardware -- the -m gpu suite, then the default bench line
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
L=gpurun_out/r2b_baseline.log
: > $L
echo "=== pytest -m gpu" >> $L
timeout 1200 python -m pytest tests -m gpu -x -q >> $L 2>&1; echo "--- exit $?" >> $L
echo "=== bench" >> $L
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/r2b_bench.json 2>> $L; echo "--- exit $?" >> $L
tail -c 600 gpurun_out/r2b_bench.json >> $L
grep -E "^===|^---|passed|failed|rror" $L | head -40
