#!/bin/bash
cd "$(dirname "$0")/.."
timeout 900 python -m pytest tests/test_conv_gpu.py tests/test_darknet_gpu.py tests/test_train_gpu.py tests/test_tiny_gpu.py tests/test_dropin_gpu.py -m gpu -q > gpurun_out/r02_pytest_fastdiv.log 2>&1; tail -3 gpurun_out/r02_pytest_fastdiv.log | cut -c1-300
timeout 600 python bench.py --workload train --no-also --no-cpu-baseline --steps 10 > gpurun_out/r02_bench_train_fd.json 2> gpurun_out/r02_bench_train_fd.err
python - <<PY
import json
j=json.loads([x for x in open("gpurun_out/r02_bench_train_fd.json") if x.startswith("{")][-1])
print("train:", round(j["ms_per_step"],2), "ms  e2e", round(j["e2e"]["ms_per_step"],2), {k:(round(v,2) if v is not None else None) for k,v in j["roofline"]["stage_ms"].items()}, j["notes"].get("rank0_per_step_ms")[1:5], j["clocks"]["sm_mhz"])
PY
timeout 600 python bench.py --workload detect --steps 3 --no-cpu-baseline > gpurun_out/r02_bench_detect_fd.json 2> gpurun_out/r02_bench_detect_fd.err; python - <<PY
import json
j=json.loads([x for x in open("gpurun_out/r02_bench_detect_fd.json") if x.startswith("{")][-1])
print("detect", round(j["value"],1), j["roofline"].get("stage_ms"), j["clocks"]["sm_mhz"])
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r02c_launches_eval_b32.csv python scratch/prof_eval.py 32 > gpurun_out/prof_eval.log 2>&1
