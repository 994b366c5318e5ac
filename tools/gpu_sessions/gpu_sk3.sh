#!/bin/bash
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
L=gpurun_out/gpu_sk3.log
: > $L
for cfg in "VLB_STREAMK=0" "VLB_SK_MAXCHUNKS=2" "VLB_SK_MAXCHUNKS=3" "VLB_SK_MAXCHUNKS=4" "VLB_SK_MAXCHUNKS=6" "VLB_SK_MAXCHUNKS=8" "VLB_SK_MAXCHUNKS=4 VLB_SK_DEBUG=2" "VLB_SK_MAXCHUNKS=4 VLB_SK_DEBUG=1"; do
  echo "=== $cfg" >> $L
  for shape in "0 6464 768 3072 256" "1 6464 768 3072 256" "1 6464 768 2304 256"; do
    env $cfg timeout 120 python tools/gemm_one.py $shape 9 >> $L 2>&1
  done
done
cat $L
