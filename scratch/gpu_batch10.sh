#!/bin/bash
cd "$(dirname "$0")/.."
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r02d_launches_train_b64.csv python scratch/prof_train.py 64 > gpurun_out/prof_train.log 2>&1
timeout 600 python bench.py --workload train --no-also --steps 10 > gpurun_out/r02_bench_train_pipe5.json 2> gpurun_out/r02_bench_train_pipe5.err; cut -c1-200 gpurun_out/r02_bench_train_pipe5.json
