#!/bin/bash
cd "$(dirname "$0")/.."
timeout 600 python -m pytest tests/test_riou_loss_gpu.py tests/test_riou_gpu.py tests/test_dropin_gpu.py -m gpu -q > gpurun_out/r02_pytest_rioul.log 2>&1; tail -12 gpurun_out/r02_pytest_rioul.log | cut -c1-300
