#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
L=gpurun_out/gpu_round9.log
: > $L
run() { echo "=== $*" >> $L; timeout "$1" "${@:2}" >> $L 2>&1; echo "--- exit $?" >> $L; }
run 600 python -m pytest tests/test_gpu_frontend.py -q -s
run 300 python tools/frontend_bench.py --images 2 --h 300 --w 500 --boxes 8 --steps 3 --torch
run 400 python tools/frontend_bench.py --steps 5 --torch --graph
grep -E "^===|^---|passed|failed|rror|parity|library|assert" $L | head -60
