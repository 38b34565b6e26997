#!/bin/bash
cd "$(dirname "$0")/.."
timeout 900 python -m pytest tests/test_bnact_gpu.py tests/test_train_gpu.py tests/test_dropin_gpu.py tests/test_tiny_gpu.py tests/test_darknet_gpu.py -m gpu -q > gpurun_out/r02_pytest_finfuse.log 2>&1; tail -3 gpurun_out/r02_pytest_finfuse.log | cut -c1-300
for b in 64 8; do
RYOLO_BENCH_PER_GPU_BATCH=$b timeout 600 python bench.py --workload train --no-also --no-cpu-baseline --steps 10 > gpurun_out/r02_bench_train_ff_pg$b.json 2> gpurun_out/r02_bench_train_ff_pg$b.err
python - <<PY
import json
j=json.loads([x for x in open("gpurun_out/r02_bench_train_ff_pg$b.json") if x.startswith("{")][-1])
print("per-gpu batch $b:", round(j["ms_per_step"],2), "ms  e2e", round(j["e2e"]["ms_per_step"],2), {k:(round(v,2) if v is not None else None) for k,v in j["roofline"]["stage_ms"].items()}, "host enqueue", j["notes"].get("rank0_host_enqueue_ms_per_step"), j["notes"].get("rank0_per_step_ms")[1:5], j["clocks"]["sm_mhz"])
PY
done
