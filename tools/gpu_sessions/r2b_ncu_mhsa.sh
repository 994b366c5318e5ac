#!/bin/bash
# round 2, session B: attention kernels alone (timing) + ncu --set full of the backward
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
L=gpurun_out/r2b_ncu_mhsa.log
: > $L
timeout 300 python tools/mhsa_one.py 64 101 12 >> $L 2>&1
timeout 300 python tools/mhsa_one.py 64 101 12 --drop 0 >> $L 2>&1
timeout 300 python tools/mhsa_one.py 64 128 12 >> $L 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:mhsa_bwd -s 6 -c 1 -f -o gpurun_out/r2b_mhsa_bwd python tools/mhsa_one.py 64 101 12 --reps 4 >> $L 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:mhsa_fwd -s 8 -c 1 -f -o gpurun_out/r2b_mhsa_fwd python tools/mhsa_one.py 64 101 12 --reps 4 >> $L 2>&1
grep -v "==PROF\|==WARN" $L | tail
