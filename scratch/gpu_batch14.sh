#!/bin/bash
cd "$(dirname "$0")/.."
timeout 900 python -m pytest tests/test_train_gpu.py tests/test_conv_gpu.py tests/test_dropin_gpu.py tests/test_tiny_gpu.py tests/test_darknet_gpu.py -m gpu -q > gpurun_out/r02_pytest_pack.log 2>&1; tail -4 gpurun_out/r02_pytest_pack.log | cut -c1-300
timeout 600 python bench.py --workload train --no-also --steps 10 > gpurun_out/r02_bench_train_pack.json 2> gpurun_out/r02_bench_train_pack.err; tail -c 300 gpurun_out/r02_bench_train_pack.err; cut -c1-200 gpurun_out/r02_bench_train_pack.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r02h_launches_train_b8.csv python scratch/prof_train.py 8 > gpurun_out/prof_train8.log 2>&1
