#!/bin/bash
# the driver's own command at N GPUs: default bench (training step + other workloads)
cd "$(dirname "$0")/.."
N=$1
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 10 --warmup 3 2> gpurun_out/r02_bench_default_n$N.err | grep '^{' > gpurun_out/r02_bench_default_n$N.json
tail -c 600 gpurun_out/r02_bench_default_n$N.err
python - <<PY
import json
j=json.load(open("gpurun_out/r02_bench_default_n$N.json"))
print("N=$N", round(j["value"],1), "img/s", round(j["ms_per_step"],2), "ms  e2e", round(j["e2e"]["ms_per_step"],2), {k:(round(v,2) if v is not None else None) for k,v in j["roofline"]["stage_ms"].items()}, j["notes"].get("rank0_per_step_ms"))
for k,w in (j.get("other_workloads") or {}).items():
    print("  ", k, w.get("value"), w.get("unit"), w.get("ms_per_step"), w.get("error"))
PY
