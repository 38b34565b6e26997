#!/bin/bash
cd "$(dirname "$0")/.."
python scratch/parity_diag64.py > gpurun_out/r02_parity_diag64.log 2>&1; tail -100 gpurun_out/r02_parity_diag64.log
for items in 16384; do
  python scratch/bn_sweep.py 64,608,608,32
  python scratch/bn_sweep.py 64,76,76,256
  python scratch/bn_sweep.py 64,19,19,1024
done > gpurun_out/r02_bn_sweep2.log 2>&1
cat gpurun_out/r02_bn_sweep2.log
timeout 300 python -m pytest tests/test_bnact_gpu.py tests/test_train_gpu.py -q > gpurun_out/r02_pytest_gpu_4.log 2>&1; tail -3 gpurun_out/r02_pytest_gpu_4.log
