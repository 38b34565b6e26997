#!/bin/bash
cd "$(dirname "$0")/.."
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r02_pytest_gpu_5.log 2>&1; tail -6 gpurun_out/r02_pytest_gpu_5.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke_b.log 2>&1; tail -2 gpurun_out/r02_smoke_b.log
timeout 1200 python bench.py > gpurun_out/r02_bench_default_b.json 2> gpurun_out/r02_bench_default_b.err; tail -c 600 gpurun_out/r02_bench_default_b.err; cut -c1-200 gpurun_out/r02_bench_default_b.json
timeout 600 python bench.py --workload riou --steps 10 > gpurun_out/r02_bench_riou_b.json 2> gpurun_out/r02_bench_riou_b.err
python scratch/bn_sweep.py 64,608,608,32 > gpurun_out/r02_bn_sweep3.log 2>&1; python scratch/bn_sweep.py 64,76,76,256 >> gpurun_out/r02_bn_sweep3.log 2>&1; cat gpurun_out/r02_bn_sweep3.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r02b_launches_train_b64.csv python scratch/prof_train.py 64 > gpurun_out/prof_train.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r02b_launches_eval_b32.csv python scratch/prof_eval.py 32 > gpurun_out/prof_eval.log 2>&1
