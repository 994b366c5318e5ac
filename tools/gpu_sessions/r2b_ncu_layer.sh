#!/bin/bash
# round 2, session B: ncu metrics of the GEMM launches of the step (current kernels: pair kernel + 128x256 kernel + grouped wgrad)
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
M=gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__throughput.avg.pct_of_peak_sustained_elapsed,dram__bytes_read.sum,dram__bytes_write.sum,lts__throughput.avg.pct_of_peak_sustained_elapsed,l1tex__throughput.avg.pct_of_peak_sustained_elapsed,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,launch__shared_mem_per_block_dynamic,sm__cycles_active.avg,sm__cycles_elapsed.max,smsp__inst_executed.sum,launch__occupancy_limit_registers,launch__occupancy_limit_shared_mem
timeout 900 ncu --metrics $M --clock-control none -k regex:"gemm_kernel|gemm_grouped|gemm_pair192" -s 130 -c 27 --csv --page raw --log-file gpurun_out/r2b_ncu_layer.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --skip-e2e --no-graph --no-gpu-eager --no-other-configs > gpurun_out/r2b_ncu_layer.log 2>&1
tail -2 gpurun_out/r2b_ncu_layer.log | cut -c1-200
echo "=== bench --dropout 0"
timeout 600 python bench.py --dropout 0 --no-cpu-baseline --no-gpu-eager --no-other-configs 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('dropout 0: ms/step %.3f value %.0f gemm frac %.3f'%(d['ms_per_step'],d['value'],d['roofline']['frac']),{k:round(v['ms_per_step'],3) for k,v in d['kernel_profile'].items()})
"
wc -l gpurun_out/r2b_ncu_layer.csv
