#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
L=gpurun_out/gpu_sk.log
: > $L
for cfg in "VLB_STREAMK=0" "VLB_STREAMK=1" "VLB_STREAMK=1 VLB_SK_MAXCHUNKS=2" "VLB_STREAMK=1 VLB_SK_MAXCHUNKS=4" "VLB_STREAMK=1 VLB_SK_DEBUG=1" "VLB_STREAMK=1 VLB_SK_DEBUG=2" "VLB_STREAMK=1 VLB_SK_DEBUG=3"; do
  echo "=== $cfg" >> $L
  for shape in "0 6464 768 3072 256" "0 6464 768 768 256" "0 6464 3072 768 256"; do
    env $cfg timeout 120 python tools/gemm_one.py $shape 9 >> $L 2>&1
  done
done
cat $L
