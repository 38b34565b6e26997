#!/bin/bash
cd "$(dirname "$0")/.."
python scratch/parity_diag.py > gpurun_out/r02_parity_diag.log 2>&1; tail -120 gpurun_out/r02_parity_diag.log
RYOLO_CONV_AROW=2 RYOLO_CONV_AROW_BO=0 timeout 600 python -m pytest tests/test_conv_gpu.py tests/test_darknet_gpu.py tests/test_train_gpu.py tests/test_tiny_gpu.py -q > gpurun_out/r02_pytest_arow2_bo0.log 2>&1; tail -3 gpurun_out/r02_pytest_arow2_bo0.log
