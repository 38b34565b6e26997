#!/bin/bash
cd "$(dirname "$0")/.."
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r02_pytest_gpu_final.log 2>&1; tail -4 gpurun_out/r02_pytest_gpu_final.log | cut -c1-300
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke_final.log 2>&1; tail -2 gpurun_out/r02_smoke_final.log
timeout 1500 python bench.py > gpurun_out/r02_bench_default_final.json 2> gpurun_out/r02_bench_default_final.err; tail -c 400 gpurun_out/r02_bench_default_final.err; cut -c1-300 gpurun_out/r02_bench_default_final.json
RYOLO_BENCH_PROFILE_RANGE=1 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r02_launches_bench_train.csv python bench.py --steps 2 --warmup 3 --no-also --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1; tail -2 gpurun_out/ncu_bench.log | cut -c1-200
