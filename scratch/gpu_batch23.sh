#!/bin/bash
cd "$(dirname "$0")/.."
timeout 900 python -m pytest tests/test_bnact_gpu.py tests/test_train_gpu.py tests/test_dropin_gpu.py tests/test_tiny_gpu.py -m gpu -q > gpurun_out/r02_pytest_rev.log 2>&1; tail -3 gpurun_out/r02_pytest_rev.log | cut -c1-300
for r in 0 3 0 3; do
RYOLO_BN_REVERSE=$r timeout 600 python bench.py --workload train --no-also --no-cpu-baseline --steps 10 > gpurun_out/r02_bench_train_rev$r.json 2> gpurun_out/r02_bench_train_rev$r.err
python - <<PY
import json
j=json.loads([x for x in open("gpurun_out/r02_bench_train_rev$r.json") if x.startswith("{")][-1])
print("REVERSE=$r:", round(j["ms_per_step"],2), "ms", {k:(round(v,2) if v is not None else None) for k,v in j["roofline"]["stage_ms"].items() if k in ("forward","loss_backward")}, j["notes"].get("rank0_per_step_ms")[1:5], j["clocks"]["sm_mhz"])
PY
done
RYOLO_BN_REVERSE=3 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r02i_launches_train_b64.csv python scratch/prof_train.py 64 > gpurun_out/prof_train.log 2>&1
