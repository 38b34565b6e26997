#!/bin/bash
# one GPU session: parity tests, whole gpu suite, A-row variant, bench lines
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_parity_gpu.py -q -s > gpurun_out/r02_parity_2.log 2>&1; tail -3 gpurun_out/r02_parity_2.log
timeout 900 python -m pytest tests -m gpu -q --deselect tests/test_parity_gpu.py > gpurun_out/r02_pytest_gpu_1.log 2>&1; tail -3 gpurun_out/r02_pytest_gpu_1.log
RYOLO_CONV_AROW=1 timeout 600 python -m pytest tests/test_conv_gpu.py tests/test_darknet_gpu.py tests/test_train_gpu.py tests/test_tiny_gpu.py -q > gpurun_out/r02_pytest_arow.log 2>&1; tail -3 gpurun_out/r02_pytest_arow.log
RYOLO_CONV_AROW=2 timeout 600 python -m pytest tests/test_conv_gpu.py tests/test_darknet_gpu.py tests/test_train_gpu.py tests/test_tiny_gpu.py -q > gpurun_out/r02_pytest_arow2.log 2>&1; tail -3 gpurun_out/r02_pytest_arow2.log
timeout 600 python bench.py --no-also --steps 5 > gpurun_out/r02_bench_train_a.json 2> gpurun_out/r02_bench_train_a.err; tail -c 600 gpurun_out/r02_bench_train_a.err
RYOLO_CONV_AROW=1 timeout 600 python bench.py --no-also --no-cpu-baseline --steps 5 > gpurun_out/r02_bench_train_arow.json 2> gpurun_out/r02_bench_train_arow.err
timeout 600 python bench.py --workload detect --steps 3 > gpurun_out/r02_bench_detect_a.json 2> gpurun_out/r02_bench_detect_a.err; tail -c 400 gpurun_out/r02_bench_detect_a.err
RYOLO_CONV_AROW=1 timeout 600 python bench.py --workload detect --steps 3 > gpurun_out/r02_bench_detect_arow.json 2> gpurun_out/r02_bench_detect_arow.err
RYOLO_CONV_AROW=2 timeout 600 python bench.py --workload detect --steps 3 > gpurun_out/r02_bench_detect_arow2.json 2> gpurun_out/r02_bench_detect_arow2.err
RYOLO_CONV_AROW=2 timeout 600 python bench.py --no-also --no-cpu-baseline --steps 5 > gpurun_out/r02_bench_train_arow2.json 2> gpurun_out/r02_bench_train_arow2.err
timeout 300 python bench.py --workload rnms --steps 5 --no-cpu-baseline > gpurun_out/r02_bench_rnms_a.json 2> gpurun_out/r02_bench_rnms_a.err
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke_a.log 2>&1; tail -2 gpurun_out/r02_smoke_a.log
for f in gpurun_out/r02_bench_*_a*.json; do echo $f; cut -c1-400 $f; done
