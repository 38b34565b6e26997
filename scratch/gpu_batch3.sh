#!/bin/bash
cd "$(dirname "$0")/.."
for items in 2048 8192 16384 65536 262144; do
  RYOLO_BN_ITEMS=$items python scratch/bn_sweep.py 64,608,608,32
  RYOLO_BN_ITEMS=$items python scratch/bn_sweep.py 64,76,76,256
done > gpurun_out/r02_bn_sweep.log 2>&1
cat gpurun_out/r02_bn_sweep.log
timeout 600 python -m pytest tests/test_parity_gpu.py -q -s -k "training" > gpurun_out/r02_parity_4.log 2>&1; grep -n "blocks \|head-layer\|passed\|failed" gpurun_out/r02_parity_4.log
timeout 300 python -m pytest tests/test_dropin_gpu.py tests/test_bnact_gpu.py -q > gpurun_out/r02_pytest_gpu_3.log 2>&1; tail -3 gpurun_out/r02_pytest_gpu_3.log
# stall reasons of the two reductions on the layer-0 tensor
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"bn_stats|bwd_reduce" -s 8 -c 2 -f -o gpurun_out/r02_prof_bn python scratch/bn_sweep.py 64,608,608,32 > gpurun_out/prof_bn.log 2>&1
