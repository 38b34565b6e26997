#!/bin/bash
cd "$(dirname "$0")/.."
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r02_pytest_gpu_final2.log 2>&1; tail -3 gpurun_out/r02_pytest_gpu_final2.log | cut -c1-300
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 10 --warmup 3 --no-also 2> gpurun_out/r02_bench_n2.err | grep '^{' > gpurun_out/r02_bench_n2.json
python - <<PY
import json
j=json.load(open("gpurun_out/r02_bench_n2.json"))
print("N=2", round(j["value"],1), "img/s", round(j["ms_per_step"],2), "ms  e2e", round(j["e2e"]["ms_per_step"],2), {k:(round(v,2) if v is not None else None) for k,v in j["roofline"]["stage_ms"].items()}, j["notes"].get("rank0_per_step_ms")[:5], j["notes"].get("rank0_host_enqueue_ms_per_step"))
PY
timeout 600 python bench.py --workload detect --steps 3 --no-cpu-baseline > gpurun_out/r02_bench_detect_final.json 2> gpurun_out/r02_bench_detect_final.err; python - <<PY
import json
j=json.loads([x for x in open("gpurun_out/r02_bench_detect_final.json") if x.startswith("{")][-1])
print("detect", round(j["value"],1), j["ms_per_step"], j["roofline"].get("stage_ms"))
PY
