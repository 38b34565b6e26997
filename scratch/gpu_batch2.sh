#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_parity_gpu.py -q -s -k "training or mini" > gpurun_out/r02_parity_3.log 2>&1; tail -3 gpurun_out/r02_parity_3.log
timeout 300 python -m pytest tests/test_dropin_gpu.py tests/test_riou_loss_gpu.py tests/test_darknet_gpu.py -q > gpurun_out/r02_pytest_gpu_2.log 2>&1; tail -3 gpurun_out/r02_pytest_gpu_2.log
RYOLO_CONV_AROW=1 RYOLO_CONV_AROW_BO=0 timeout 300 python -m pytest tests/test_conv_gpu.py -q -k vs_torch > gpurun_out/r02_pytest_arow_bo0.log 2>&1; tail -3 gpurun_out/r02_pytest_arow_bo0.log
# launch lists (eager launches, profiler range = one step / one forward)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r02_launches_train_b64.csv python scratch/prof_train.py 64 > gpurun_out/prof_train.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r02_launches_eval_b32.csv python scratch/prof_eval.py 32 > gpurun_out/prof_eval.log 2>&1
# thin layer (32 -> 64, 3x3 @304^2) and one body layer, full sets
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -c 1 -f -o gpurun_out/r02_prof_thin python scratch/one_layer.py 32,64,3,304 2 > gpurun_out/prof_thin.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -c 1 -f -o gpurun_out/r02_prof_thin_s2d python scratch/one_layer.py 128,64,2,304 2 > gpurun_out/prof_thin2.log 2>&1
# default bench line (all workloads, incl. parity-precision training step at batch 64)
timeout 1200 python bench.py > gpurun_out/r02_bench_default_a.json 2> gpurun_out/r02_bench_default_a.err; tail -c 500 gpurun_out/r02_bench_default_a.err; cut -c1-300 gpurun_out/r02_bench_default_a.json
