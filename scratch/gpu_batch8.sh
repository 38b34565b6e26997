#!/bin/bash
cd "$(dirname "$0")/.."
timeout 600 python -m pytest tests/test_riou_gpu.py tests/test_dropin_gpu.py tests/test_riou_loss_gpu.py -m gpu -q > gpurun_out/r02_pytest_riou.log 2>&1; tail -8 gpurun_out/r02_pytest_riou.log
timeout 300 python scratch/prof_riou.py 10 2>&1 | tail -2
timeout 600 ncu --set full --import-source on --clock-control none -k regex:riou_pairwise -c 1 -o gpurun_out/r02_riou_v5 -f python scratch/prof_riou.py 1 > gpurun_out/ncu_riou.log 2>&1; tail -2 gpurun_out/ncu_riou.log
timeout 600 python bench.py --workload riou --steps 10 > gpurun_out/r02_bench_riou_d.json 2> gpurun_out/r02_bench_riou_d.err; cut -c1-400 gpurun_out/r02_bench_riou_d.json
