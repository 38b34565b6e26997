#!/bin/bash
cd "$(dirname "$0")/.."
N=$1
run() { tag=$1; shift; timeout 600 env "$@" python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 10 --warmup 3 --no-also 2> gpurun_out/r02_n${N}_$tag.err | grep '^{' > gpurun_out/r02_n${N}_$tag.json; python - <<PY
import json
j=json.load(open("gpurun_out/r02_n${N}_$tag.json"))
print("$tag", round(j["value"],1), "img/s", round(j["ms_per_step"],2), "ms  e2e", round(j["e2e"]["ms_per_step"],2), {k:(round(v,2) if v is not None else None) for k,v in j["roofline"]["stage_ms"].items()}, j["notes"].get("rank0_per_step_ms")[1:5], j["notes"].get("rank0_host_enqueue_ms_per_step"))
PY
}
run default X=1
run r8_ch8 RYOLO_DDP_RESERVED_SMS=8 NCCL_MAX_NCHANNELS=8
run r16_ch16 RYOLO_DDP_RESERVED_SMS=16 NCCL_MAX_NCHANNELS=16
run b64 RYOLO_DDP_BUCKET_MB=64
run b16 RYOLO_DDP_BUCKET_MB=16
run r4_ch4 RYOLO_DDP_RESERVED_SMS=4 NCCL_MAX_NCHANNELS=4
