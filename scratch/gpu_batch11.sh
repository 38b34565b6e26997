#!/bin/bash
cd "$(dirname "$0")/.."
timeout 900 python -m pytest tests/test_loss_gpu.py tests/test_train_gpu.py tests/test_dropin_gpu.py tests/test_tiny_gpu.py tests/test_parity_gpu.py -m gpu -q -x > gpurun_out/r02_pytest_loss.log 2>&1; tail -12 gpurun_out/r02_pytest_loss.log
timeout 600 python bench.py --workload train --no-also --steps 10 > gpurun_out/r02_bench_train_loss.json 2> gpurun_out/r02_bench_train_loss.err; tail -c 300 gpurun_out/r02_bench_train_loss.err; cut -c1-200 gpurun_out/r02_bench_train_loss.json
