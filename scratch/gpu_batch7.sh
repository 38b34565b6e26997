#!/bin/bash
cd "$(dirname "$0")/.."
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r02_pytest_gpu_7.log 2>&1; tail -6 gpurun_out/r02_pytest_gpu_7.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke_c.log 2>&1; tail -2 gpurun_out/r02_smoke_c.log
timeout 1200 python bench.py > gpurun_out/r02_bench_default_c.json 2> gpurun_out/r02_bench_default_c.err; tail -c 600 gpurun_out/r02_bench_default_c.err; cut -c1-300 gpurun_out/r02_bench_default_c.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r02c_launches_train_b64.csv python scratch/prof_train.py 64 > gpurun_out/prof_train.log 2>&1
