#!/bin/bash
# bench + ncu launch list (+ optional full capture of the GEMM kernels).  Usage: tools/gpu_bench.sh [full]
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
L=gpurun_out/gpu_bench.log
: > $L
run() { echo "=== $*" >> $L; timeout "$1" "${@:2}" >> $L 2>&1; echo "--- exit $?" >> $L; }
run 600 python -m pytest tests/test_gpu_kernels.py -q -k "gemm" -x
run 600 python -m pytest tests/test_gpu_modules.py -q -x
run 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline
run 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --skip-e2e
if [ "$1" = "full" ]; then
  run 900 ncu --set full --clock-control none --import-source on -k regex:gemm_kernel -s 144 -c 36 -f -o gpurun_out/prof_gemm \
      python bench.py --steps 1 --warmup 3 --no-cpu-baseline --skip-e2e
fi
grep -E "^===|^---|passed|failed|rror|metric" $L | cut -c1-1500 | head -40
python tools/summarize_launches.py gpurun_out/launches.csv | tee gpurun_out/launches_summary.txt | head -60
