#!/bin/bash
cd "$(dirname "$0")/.."
timeout 900 python -m pytest tests/test_bnact_gpu.py tests/test_train_gpu.py tests/test_dropin_gpu.py tests/test_tiny_gpu.py -m gpu -q -x > gpurun_out/r02_pytest_bnpipe.log 2>&1; tail -8 gpurun_out/r02_pytest_bnpipe.log
for sz in 64,608,608,32 64,304,304,64 64,152,152,128 64,76,76,256 64,38,38,512 64,19,19,1024 64,76,76,128; do
  RYOLO_BN_PIPE=0 python scratch/bn_sweep.py $sz 2>&1 | tail -1 | sed 's/^/direct /'
  RYOLO_BN_PIPE=7 python scratch/bn_sweep.py $sz 2>&1 | tail -1 | sed 's/^/pipe   /'
done > gpurun_out/r02_bn_sweep_pipe.txt 2>&1; cat gpurun_out/r02_bn_sweep_pipe.txt
timeout 600 python bench.py --workload train --no-also --steps 10 > gpurun_out/r02_bench_train_pipe.json 2> gpurun_out/r02_bench_train_pipe.err; tail -c 300 gpurun_out/r02_bench_train_pipe.err; cut -c1-200 gpurun_out/r02_bench_train_pipe.json
RYOLO_BN_PIPE=0 timeout 600 python bench.py --workload train --no-also --steps 10 > gpurun_out/r02_bench_train_nopipe.json 2> /dev/null; cut -c1-200 gpurun_out/r02_bench_train_nopipe.json
