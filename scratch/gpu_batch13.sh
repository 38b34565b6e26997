#!/bin/bash
cd "$(dirname "$0")/.."
timeout 900 python -m pytest tests/test_loss_gpu.py tests/test_train_gpu.py tests/test_dropin_gpu.py tests/test_tiny_gpu.py -m gpu -q > gpurun_out/r02_pytest_loss3.log 2>&1; tail -5 gpurun_out/r02_pytest_loss3.log | cut -c1-300
timeout 600 python bench.py --workload train --no-also --steps 10 > gpurun_out/r02_bench_train_loss3.json 2> gpurun_out/r02_bench_train_loss3.err; tail -c 300 gpurun_out/r02_bench_train_loss3.err; cut -c1-200 gpurun_out/r02_bench_train_loss3.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r02g_launches_train_b64.csv python scratch/prof_train.py 64 > gpurun_out/prof_train.log 2>&1
