#!/bin/bash
# scratch/gpu_batch_n.sh <N>: DDP tests + the default training bench line at N GPUs
cd "$(dirname "$0")/.."
N=$1
timeout 600 python -m pytest tests/test_ddp_gpu.py -m gpu -q > gpurun_out/r02_pytest_ddp_n$N.log 2>&1; tail -3 gpurun_out/r02_pytest_ddp_n$N.log | cut -c1-200
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 10 --warmup 3 --no-also 2> gpurun_out/r02_bench_n$N.err | grep '^{' > gpurun_out/r02_bench_n$N.json
tail -c 400 gpurun_out/r02_bench_n$N.err
python - <<PY
import json
j=json.load(open("gpurun_out/r02_bench_n$N.json"))
print("N=$N", round(j["value"],1), "img/s", round(j["ms_per_step"],2), "ms  e2e", round(j["e2e"]["ms_per_step"],2), {k:(round(v,2) if v is not None else None) for k,v in j["roofline"]["stage_ms"].items()}, j["notes"].get("rank0_per_step_ms"))
PY
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus $N --steps 10 --warmup 3 --workload riou --no-also 2> gpurun_out/r02_bench_riou_n$N.err | grep '^{' > gpurun_out/r02_bench_riou_n$N.json; cut -c1-160 gpurun_out/r02_bench_riou_n$N.json
