#!/bin/bash
cd "$(dirname "$0")/.."
run() { tag=$1; shift; timeout 600 env "$@" python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 8 --warmup 3 --no-also $EXTRA 2> gpurun_out/r02_n2_$tag.err | grep '^{' > gpurun_out/r02_n2_$tag.json; python - <<PY
import json
j=json.load(open("gpurun_out/r02_n2_$tag.json"))
print("$tag", round(j["value"],1), "img/s", round(j["ms_per_step"],2), "ms  e2e", round(j["e2e"]["ms_per_step"],2), j["roofline"]["stage_ms"], j["notes"].get("rank0_per_step_ms"), j["notes"].get("reserved_sms"))
PY
}
EXTRA="" run graph_r8 RYOLO_DDP_RESERVED_SMS=8
EXTRA="" run graph_r0 RYOLO_DDP_RESERVED_SMS=0
EXTRA="--no-graph" run eager_r0 RYOLO_DDP_RESERVED_SMS=0
EXTRA="--no-graph" run eager_r8 RYOLO_DDP_RESERVED_SMS=8
EXTRA="" run graph_r0_ch16 RYOLO_DDP_RESERVED_SMS=0 NCCL_MAX_NCHANNELS=16
timeout 300 python -m pytest tests/test_train_gpu.py tests/test_bnact_gpu.py tests/test_dropin_gpu.py -q 2>&1 | tail -3
