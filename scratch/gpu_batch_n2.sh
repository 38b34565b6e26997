#!/bin/bash
cd "$(dirname "$0")/.."
nvidia-smi -L
timeout 600 python -m pytest tests/test_ddp_gpu.py -q -s > gpurun_out/r02_pytest_ddp.log 2>&1; tail -5 gpurun_out/r02_pytest_ddp.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r02_bench_n2.json 2> gpurun_out/r02_bench_n2.err; tail -c 1500 gpurun_out/r02_bench_n2.err; cut -c1-600 gpurun_out/r02_bench_n2.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 5 --warmup 3 --no-also --no-graph > gpurun_out/r02_bench_n2_nograph.json 2> gpurun_out/r02_bench_n2_nograph.err; cut -c1-300 gpurun_out/r02_bench_n2_nograph.json
