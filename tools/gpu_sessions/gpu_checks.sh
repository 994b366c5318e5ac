#!/bin/bash
# Run the GPU parity tests section by section (separate processes: a trapped kernel poisons its CUDA context),
# then smoke() and a short bench.  Logs land in gpurun_out/.
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
L=gpurun_out/gpu_checks.log
: > $L
run() { echo "=== $*" >> $L; timeout "$1" "${@:2}" >> $L 2>&1; echo "--- exit $?" >> $L; }
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv >> $L
run 900 python -m pytest tests/test_gpu_kernels.py -q -k "gemm" -x
run 300 python -m pytest tests/test_gpu_kernels.py -q -k "layernorm"
run 300 python -m pytest tests/test_gpu_kernels.py -q -k "mhsa"
run 300 python -m pytest tests/test_gpu_kernels.py -q -k "roi"
run 900 python -m pytest tests/test_gpu_modules.py -q
run 300 python -c "import __graft_entry__ as g; g.smoke()"
run 300 python tools/report_parity.py
run 200 python tools/cpu_threads_probe.py
run 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline
grep -E "^===|^---|passed|failed|error|Error|smoke|metric|^L=|threads|cpu_count" $L | head -80
