#!/usr/bin/env python
"""bench.py -- one JSON line per run (contract in the task statement / DESIGN.md "Measurement").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload riou|rnms|detect] [--impl ours|reference]

Default (N=1): BASELINE.json configs[1] -- rotated IoU, 10k x 10k random (cx,cy,w,h,theta) boxes -> Mpairs/s.
A "step" is one pass of the hot path over one batch of synthetic input.  `value` is measured with inputs
resident in HBM; `e2e` goes through the public Python API with pinned HOST buffers (H2D + D2H inside the timed
region).  `roofline` is for the dominant kernel (CUDA events on the launching stream, algorithmic bytes from
SURVEY.md 8d); `cpu_baseline` is the reference's own kernel code compiled for the host (oracle/_ref, kind
"reference") or the oracle port, on a bounded sample.  `--impl reference` times only that CPU path.
Multi-GPU (torchrun): the path shards by independent row blocks / images, no data-path collective, weak scaling;
time = max over ranks."""
import argparse
import ctypes
import json
import math
import os
import subprocess
import sys
import threading
import time

REPO = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "tests"))


def peaks():
    p = os.path.join(REPO, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm_gbs": float(d["hbm_gbs"]), "bf16_tflops": float(d["bf16_tflops"]),
                "bf16_tflops_sustained": float(d.get("bf16_tflops_sustained", d["bf16_tflops"])), "src": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "src": "fallback"}


WORKLOADS = {   # config.workload of both arms (ours / --impl reference)
    "riou": "rotated IoU, N=M=10000 random (cx,cy,w,h,theta) boxes on a 608^2 canvas (BASELINE configs[1])",
    "rnms": "rotated NMS, 20000 boxes/image, 1 class, IoU thr 0.5 (BASELINE configs[2])",
    "train": "Darknet-53 (cfg/yolov3.cfg graph) training step: fwd (batch-stat BN) + compute_loss + bwd + SGD, "
             "608x608 synthetic (BASELINE configs[3])",
    "detect": "detect e2e: Darknet-53 (cfg/yolov3.cfg graph) eval forward + YOLO decode + conf filter + top-20000 + "
              "rotated NMS thr 0.5, 608x608 (BASELINE configs[4] shape)",
}


class ClockSampler:
    """SM clock / throttle reasons DURING the timed region (B200_PROFILING.md clocks line).  NVML is polled from a thread
    every 2 ms (an `nvidia-smi -lms` child needs ~0.3 s to produce its first line -- longer than most timed regions
    here); nvidia-smi is the fallback when the NVML binding is missing."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")
    BITS = {0x4: "sw_power_cap", 0x8: "hw_slowdown", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown"}

    def __init__(self, gpu_index):
        self.rows, self.proc, self.idx = [], None, gpu_index
        self.nvml, self.samples, self.reasons, self.stop_flag, self.mx = None, [], set(), False, None

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            h = pynvml.nvmlDeviceGetHandleByIndex(self.idx)
            self.mx = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))
            get_reasons = getattr(pynvml, "nvmlDeviceGetCurrentClocksEventReasons", None) or \
                getattr(pynvml, "nvmlDeviceGetCurrentClocksThrottleReasons")

            def poll():
                while not self.stop_flag:
                    try:
                        self.samples.append(float(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)))
                        r = int(get_reasons(h))
                        for bit, name in self.BITS.items():
                            if r & bit:
                                self.reasons.add(name)
                    except Exception:
                        pass
                    time.sleep(0.002)
            self.nvml = threading.Thread(target=poll, daemon=True)
            self.nvml.start()
            return
        except Exception:
            self.nvml = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.nvml is not None:
            self.stop_flag = True
            self.nvml.join(timeout=1)
            sm = sorted(self.samples)
            return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": self.mx, "reasons": sorted(self.reasons),
                    "samples": len(sm), "source": "nvml, 2 ms period"}
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            pass
        sm = sorted(float(r[1]) for r in self.rows if len(r) >= 8 and r[1].replace(".", "").isdigit())
        mx = [float(r[2]) for r in self.rows if len(r) >= 8 and r[2].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) >= 8:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx[0] if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "source": "nvidia-smi -lms 100"}


# ---------------------------------------------------------------------------------------------------------------
# CPU legs (the only place bench.py touches oracle/)
# ---------------------------------------------------------------------------------------------------------------
def cpu_threads():
    return max(1, len(os.sched_getaffinity(0)))


def cpu_riou(sample_rows, steps=1):
    """reference IoU code on the host cores over `sample_rows` x 10k pairs of the config-2 workload"""
    import numpy as np
    import helpers
    a = np.concatenate([helpers.gen_boxes(10000, 0).numpy(), np.zeros((10000, 1), np.float32)], 1)[:sample_rows]
    b = np.concatenate([helpers.gen_boxes(10000, 1).numpy(), np.zeros((10000, 1), np.float32)], 1)
    a, b = np.ascontiguousarray(a), np.ascontiguousarray(b)
    out = np.empty((sample_rows, 10000), np.float32)
    ref = helpers.ref_lib("host")
    th = cpu_threads()
    times = []
    for _ in range(steps):
        t0 = time.perf_counter()
        if ref is not None:
            ref.ref_host_iou_pairwise(helpers.P(a), sample_rows, helpers.P(b), 10000, 6, helpers.P(out), th)
            kind, cores = "reference", th
        else:
            helpers.oracle().orc_skew_iou_pairwise(helpers.P(a), sample_rows, 6, helpers.P(b), 10000, 6, 0, helpers.P(out))
            kind, cores = "port", 1
        times.append(time.perf_counter() - t0)
    pairs = sample_rows * 10000
    what = ("reference device IoU code (rotate_polygon_nms_kernel.cu:22-260) as host C++, OpenMP" if kind == "reference"
            else "oracle float64 clip (orc_skew_iou), scalar")
    return {"value": pairs / min(times) / 1e6, "unit": "Mpairs/s", "cores": cores, "kind": kind,
            "sample": "%d x 10000 pairs of the 10k x 10k workload; %s" % (sample_rows, what)}, times


def cpu_rnms(n, steps=1):
    import numpy as np
    import helpers
    dets = helpers.gen_dets(n, 2).numpy()
    keep = np.empty(n, np.int64)
    ref = helpers.ref_lib("host")
    th = cpu_threads()
    times = []
    for _ in range(steps):
        t0 = time.perf_counter()
        if ref is not None:
            k = ref.ref_host_rnms(helpers.P(dets), n, ctypes.c_float(0.5), keep.ctypes.data_as(helpers.I64P), th, None, None)
            kind, cores = "reference", th
        else:
            k = len(helpers.orc_rnms(dets, 0.5))
            kind, cores = "port", 1
        times.append(time.perf_counter() - t0)
    return {"value": n / min(times), "unit": "boxes/s", "cores": cores, "kind": kind,
            "sample": "%d boxes (config-3 generator), thr 0.5, K=%d; upper-triangle IoU + reference serial scan" % (n, k)}, times


def torch_port_forward(model, x, train):
    """PyTorch restatement of the reference Darknet.forward graph walk (model/models.py:244-298) on the modules of
    `model` (same names / shapes as the reference's): the CPU baseline of the conv stacks."""
    import torch
    import torch.nn.functional as F
    outs, heads = [], []
    for i, (d, mod) in enumerate(zip(model.module_defs, model.module_list)):
        t = d["type"]
        if t == "convolutional":
            k = mod.Conv2d.weight.shape[-1]
            y = F.conv2d(x, mod.Conv2d.weight, mod.Conv2d.bias, stride=int(d["stride"]), padding=(k - 1) // 2)
            if hasattr(mod, "BatchNorm2d"):
                bn = mod.BatchNorm2d
                y = F.batch_norm(y, None if train else bn.running_mean, None if train else bn.running_var, bn.weight, bn.bias,
                                 training=train, eps=bn.eps)
            if hasattr(mod, "activation"):
                y = F.prelu(y, mod.activation.weight)
            x = y
            nxt = model.module_defs[i + 1]["type"] if i + 1 < len(model.module_defs) else None
            if nxt == "yolo" or (not model.yolo_layers and not hasattr(mod, "BatchNorm2d") and not hasattr(mod, "activation")):
                heads.append(y)
        elif t == "shortcut":
            x = x + outs[i + int(d["from"])]
        elif t == "upsample":
            x = F.interpolate(x, scale_factor=2, mode="nearest")
        elif t == "maxpool":
            x = mod(x)          # nn.MaxPool2d / Sequential(ZeroPad2d, MaxPool2d) exactly as create_modules builds them
        elif t == "route":
            ls = [l if l > 0 else i + l for l in (int(v) for v in d["layers"].split(","))]
            x = torch.cat([outs[l] for l in ls], 1) if len(ls) > 1 else outs[ls[0]]
        outs.append(x)
    if not model.yolo_layers:
        return heads
    res = []
    for hd, yi in zip(heads, model.yolo_layers):
        layer = model.module_list[yi]
        res.append(hd.view(hd.shape[0], layer.na, layer.nc + 6, hd.shape[2], hd.shape[3]).permute(0, 1, 3, 4, 2).contiguous())
    return res


TRAIN_HYP = {"giou": 0.1, "cls": 27.76, "cls_pw": 1.0, "obj": 20.35, "obj_pw": 1.0, "iou_t": 0.5, "ang_t": 3.1415926 / 12,
             "reg": 1.0, "context_factor": 1.0}    # cfg/hyp_template.py with cls_pw = obj_pw = 1 (arc 'default', train.py:43-45)


def make_targets(n_img, seed):
    """3 boxes / image: (img, 0, cx,cy~U(.1,.9), w~U(.05,.35), h = w/U(3,9), theta~U(-pi/2,pi/2))"""
    import torch
    g = torch.Generator().manual_seed(seed)
    nt = 3 * n_img
    t = torch.zeros(nt, 7)
    t[:, 0] = torch.arange(n_img).repeat_interleave(3).float()
    t[:, 2:4] = 0.1 + 0.8 * torch.rand(nt, 2, generator=g)
    t[:, 4] = 0.05 + 0.30 * torch.rand(nt, generator=g)
    t[:, 5] = t[:, 4] / (3 + 6 * torch.rand(nt, generator=g))
    t[:, 6] = (torch.rand(nt, generator=g) - 0.5) * math.pi
    return t


def cpu_train(steps=1, batch=2):
    """fwd + loss + bwd of the same graph with stock PyTorch CPU kernels (what the reference's nn.Modules would run)"""
    import torch
    import helpers
    import rotate_yolov3_b200 as pkg
    from rotate_yolov3_b200 import cfgs
    from rotate_yolov3_b200.loss import compute_loss
    model = pkg.Darknet(cfgs.yolov3_cfg(), dict(TRAIN_HYP), arc="default")
    helpers.init_darknet_weights(model, seed=1)
    model.nc, model.hyp = 1, dict(TRAIN_HYP)
    x = torch.rand(batch, 3, 608, 608)
    tg = make_targets(batch, 5)
    for yi in model.yolo_layers:
        pass
    times = []
    for _ in range(steps):
        t0 = time.perf_counter()
        ps = torch_port_forward(model, x, True)
        for p, yi in zip(ps, model.yolo_layers):
            layer = model.module_list[yi]
            if (layer.nx, layer.ny) != (p.shape[3], p.shape[2]):
                layer.create_grids((608, 608), (p.shape[3], p.shape[2]), "cpu", torch.float32)
        loss, _ = compute_loss(ps, tg.clone(), model, model.hyp)
        loss.backward()
        times.append(time.perf_counter() - t0)
        for p in model.parameters():
            p.grad = None
    return {"value": batch / min(times), "unit": "images/s", "cores": torch.get_num_threads(), "kind": "port",
            "sample": "%d images 608x608, fwd + compute_loss + bwd, stock PyTorch CPU kernels on the same graph "
                      "(the reference's own model cannot be imported on the GPU box)" % batch}, times


# ---------------------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--workload", default="riou", choices=["riou", "rnms", "detect", "train"])
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-also", action="store_true", help="skip the secondary workloads summary in the default run")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    K, W = args.steps, max(args.warmup, 3) if args.impl == "ours" else args.warmup

    metric = {"riou": ("rotated-IoU Mpairs/sec", "Mpairs/s"), "rnms": ("RNMS boxes/sec", "boxes/s"),
              "detect": ("608x608 images/sec", "images/s"), "train": ("608x608 images/sec", "images/s")}[args.workload]

    if args.impl == "reference":
        # reference arm: the reference's own CPU implementation of the path on the host cores, rank 0 only
        if rank != 0:
            return
        if args.workload == "riou":
            cb, times = cpu_riou(sample_rows=10000, steps=max(1, min(K, 3)) + min(W, 1))
            sample = "the full 1e8 pairs per step"
        elif args.workload == "rnms":
            cb, times = cpu_rnms(20000, steps=max(1, min(K, 2)))
            sample = "the full 20000 boxes per step"
        elif args.workload == "train":
            cb, times = cpu_train(steps=max(1, min(K, 2)), batch=2)
            sample = "bounded sample of 2 images per step (stock PyTorch CPU kernels on the same graph)"
        else:
            print(json.dumps({"impl": "reference", "unavailable": "detect workload has no CPU reference arm yet"}))
            return
        print(json.dumps({"impl": "reference", "metric": metric[0], "value": cb["value"], "unit": metric[1],
                          "n_gpus": args.gpus, "steps": K, "warmup": W, "ms_per_step": 1e3 * min(times),
                          "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                          "dtype": "f32", "data": "synthetic",
                          "config": {"workload": WORKLOADS[args.workload], "sample": sample},
                          "cpu_baseline": cb,
                          "e2e": {"value": cb["value"], "unit": metric[1], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))
        return

    import torch
    import torch.distributed as dist
    import helpers
    import rotate_yolov3_b200 as pkg
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    pk = peaks()
    lib = pkg._lib.lib

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------- workload setup ----------------
    if args.workload == "riou":
        n = m = 10000
        a_h = helpers.gen_boxes(n, 100 * rank).pin_memory()   # rank r owns row block r of a (world*10k) x 10k problem
        b_h = helpers.gen_boxes(m, 1).pin_memory()
        a, b = a_h.to(dev), b_h.to(dev)
        outs = [torch.empty((n, m), dtype=torch.float32, device=dev) for _ in range(2)]
        out_h = torch.empty((n, m), dtype=torch.float32).pin_memory()
        units = n * m
        alg_bytes = 20 * (n + m) + 4 * n * m      # 400 400 000 B per call
        launches_per_step = 1

        def step(i):
            pkg.rotated_iou_matrix(a, b, out=outs[i & 1])

        def step_e2e(i):
            ad = a_h.to(dev, non_blocking=True)
            bd = b_h.to(dev, non_blocking=True)
            pkg.rotated_iou_matrix(ad, bd, out=outs[i & 1])
            out_h.copy_(outs[i & 1], non_blocking=True)
            torch.cuda.current_stream().synchronize()
        h2d, d2h = (n + m) * 5 * 4, n * m * 4
        cfg = {"workload": WORKLOADS["riou"],
               "mode": "iou", "sharding": "row blocks of A per rank, B replicated, no collective",
               "l2": "each step streams a 400 MB output (> 126 MB L2) into alternating buffers"}
        scale = 1e-6
    elif args.workload == "rnms":
        n = 20000
        d_h = helpers.gen_dets(n, 2 + 100 * rank).pin_memory()
        d = d_h.to(dev)
        units = n
        cbk = (n + 63) // 64
        alg_bytes = 24 * n + 8 * n * cbk * 2       # boxes in + mask write + mask read (reference algorithm)
        launches_per_step = 8
        flush = torch.empty(192 << 20, dtype=torch.uint8, device=dev)
        keep_h = torch.empty(n, dtype=torch.long).pin_memory()

        def step(i):
            pkg.r_nms(d, 0.5)

        def step_e2e(i):
            dd = d_h.to(dev, non_blocking=True)
            k = pkg.r_nms(dd, 0.5)
            keep_h[:len(k)].copy_(k)
        h2d, d2h = n * 6 * 4, 8 * 8666
        cfg = {"workload": WORKLOADS["rnms"],
               "sharding": "one image per rank (replicas), no collective",
               "l2": "50 MB mask rewritten every step; 192 MB flush buffer written between timed steps"}
        scale = 1.0
    elif args.workload == "train":
        # BASELINE configs[3]: Darknet-53 training step, global batch 64 x 608 x 608 synthetic, compute_loss of the
        # reference (no rotated IoU in it), SGD nesterov; images shard across ranks, per-replica BN, ONE
        # NCCL all-reduce over the flattened gradients per step.
        from rotate_yolov3_b200 import cfgs, parallel
        from rotate_yolov3_b200.loss import compute_loss
        per_gpu = max(1, 64 // world)
        model = pkg.Darknet(cfgs.yolov3_cfg(), dict(TRAIN_HYP), arc="default")
        helpers.init_darknet_weights(model, seed=1)
        model.nc, model.hyp = 1, dict(TRAIN_HYP)
        model = model.to(dev).train()
        pg_w = [p for n, p in model.named_parameters() if "Conv2d.weight" in n]
        pg_o = [p for n, p in model.named_parameters() if "Conv2d.weight" not in n]
        opt = torch.optim.SGD([{"params": pg_o}, {"params": pg_w, "weight_decay": 4.569e-4}], lr=1e-4, momentum=0.97,
                              nesterov=True)                       # train.py:70-82 param groups, cfg/hyp_template.py
        x_h = torch.rand(per_gpu, 3, 608, 608, generator=torch.Generator().manual_seed(rank)).pin_memory()
        tg_h = make_targets(per_gpu, 100 + rank).pin_memory()
        x, tg = x_h.to(dev), tg_h.to(dev)
        params = [p for p in model.parameters()]
        units = per_gpu
        alg_bytes = None
        launches_per_step = 700
        loss_h = torch.empty(1).pin_memory()
        stage_ms = {}

        def train_step(xd, td, timers=None):
            opt.zero_grad(set_to_none=True)
            if timers:
                timers[0].record()
            ps = model(xd)
            if timers:
                timers[1].record()
            loss, items = compute_loss(ps, td.clone(), model, model.hyp)
            loss.backward()
            if timers:
                timers[2].record()
            parallel.allreduce_gradients(params)      # one flat NCCL all-reduce (no-op at world size 1)
            opt.step()
            if timers:
                timers[3].record()
            return loss

        def step(i):
            train_step(x, tg)

        def step_e2e(i):
            xd = x_h.to(dev, non_blocking=True)
            td = tg_h.to(dev, non_blocking=True)
            loss = train_step(xd, td)
            loss_h.copy_(loss.detach(), non_blocking=True)
            torch.cuda.current_stream().synchronize()
        h2d, d2h = per_gpu * 3 * 608 * 608 * 4 + tg_h.numel() * 4, 4
        cfg = {"workload": WORKLOADS["train"],
               "global_batch": per_gpu * world, "per_gpu_batch": per_gpu, "precision": "bf16 operands/activations, fp32 "
               "accumulate and parameter gradients", "parallelism": "dp%d" % world,
               "collective": "one flat NCCL all-reduce of 62.4M fp32 gradients per step (after backward, not overlapped)",
               "l2": "activations of one step (tens of GB) exceed the 126 MB L2"}
        scale = 1.0
    else:
        # BASELINE configs[4] shape: eval forward (conv stacks + decode) -> conf filter -> per-image top-20000 -> RNMS.
        # Per-GPU batch = 32 / world (images shard across ranks, no collective); random-init Darknet-53 with randomised
        # BN statistics / PReLU slopes.
        from rotate_yolov3_b200 import cfgs
        from rotate_yolov3_b200.nms import nms_filter_async, r_nms_async
        per_gpu = max(1, 32 // world)
        model = pkg.Darknet(cfgs.yolov3_cfg(), {"context_factor": 1.0}, arc="default")
        helpers.init_darknet_weights(model, seed=1)
        model = model.to(dev).eval()
        model.use_cuda_graph = True      # the 79 launches of one forward replayed as one CUDA graph
        x_h = torch.rand(per_gpu, 3, 608, 608, generator=torch.Generator().manual_seed(rank)).pin_memory()
        x = x_h.to(dev)
        CAP, CONF, THR = 20000, 0.5, 0.5
        units = per_gpu
        alg_bytes = None
        launches_per_step = 80
        streams = [torch.cuda.Stream(device=dev) for _ in range(4)]
        stage_ms = {}

        def run(xd, timers=None):
            with torch.no_grad():
                if timers:
                    timers[0].record()
                io, _ = model(xd)
                if timers:
                    timers[1].record()
                dets = []
                for i in range(per_gpu):
                    cand, _num = nms_filter_async(io[i], CONF, 2.0, 300000)
                    top = torch.topk(cand[:, 5], CAP).indices
                    dets.append(cand[top][:, :6].contiguous())
                if timers:
                    timers[2].record()
                ev0 = torch.cuda.Event()
                ev0.record()
                outs = []
                for i in range(per_gpu):
                    st = streams[i % len(streams)]
                    st.wait_event(ev0)
                    with torch.cuda.stream(st):
                        outs.append(r_nms_async(dets[i], THR))
                for st in streams:
                    torch.cuda.current_stream().wait_stream(st)
                if timers:
                    timers[3].record()
                counts = torch.cat([o[1] for o in outs])
            return counts, outs, dets

        def step(i):
            run(x)

        counts_h = torch.empty(per_gpu, dtype=torch.int32).pin_memory()

        def step_e2e(i):
            xd = x_h.to(dev, non_blocking=True)
            counts, outs, dets = run(xd)
            counts_h.copy_(counts, non_blocking=True)
            torch.cuda.current_stream().synchronize()
        h2d, d2h = per_gpu * 3 * 608 * 608 * 4, per_gpu * 4
        cfg = {"workload": WORKLOADS["detect"],
               "global_batch": per_gpu * world, "per_gpu_batch": per_gpu, "precision": "bf16 operands, fp32 accumulate",
               "sharding": "images per rank, no collective",
               "l2": "activations of one forward (~8 GB at batch 32) exceed the 126 MB L2"}
        scale = 1.0
    # ---------------- device-resident timing ----------------
    for i in range(W):
        step(i)
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    sampler = ClockSampler(local_rank)
    barrier()
    if rank == 0:
        sampler.start()
    l0 = lib.ryolo_launch_count()
    t_start = torch.cuda.Event(enable_timing=True)
    t_end = torch.cuda.Event(enable_timing=True)
    t_start.record()
    flush_ms = 0.0
    for i in range(K):
        if args.workload == "rnms":
            flush.fill_(i & 0xFF)   # L2 flush between timed iterations (not counted: per-step events below)
        ev[i][0].record()
        step(i)
        ev[i][1].record()
    t_end.record()
    barrier()
    launches = lib.ryolo_launch_count() - l0
    clocks = sampler.stop() if rank == 0 else None
    per_step = [s.elapsed_time(e) for s, e in ev]
    tot_ms = sum(per_step)
    t = torch.tensor([tot_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    tot_ms = float(t.item())
    ms_per_step = tot_ms / K
    value = units * world / (ms_per_step * 1e-3) * scale

    # ---------------- end-to-end through the public API, host buffers ----------------
    for i in range(3):
        step_e2e(i)
    barrier()
    e0 = torch.cuda.Event(enable_timing=True)
    e1 = torch.cuda.Event(enable_timing=True)
    Ke = max(3, K // 2)
    e0.record()
    for i in range(Ke):
        step_e2e(i)
    e1.record()
    barrier()
    te = torch.tensor([e0.elapsed_time(e1) / Ke], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_val = units * world / (float(te.item()) * 1e-3) * scale

    # one more step with stage timers.  EVERY rank runs it: the training step contains the gradient all-reduce, so a
    # rank-0-only step would wait for peers that have already left.
    if args.workload == "train":
        tm = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        train_step(x, tg, tm)
        torch.cuda.synchronize()
        barrier()
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---------------- roofline of the dominant kernel ----------------
    if args.workload == "train":
        stage_ms = {"forward": tm[0].elapsed_time(tm[1]), "loss_backward": tm[1].elapsed_time(tm[2]),
                    "allreduce_sgd": tm[2].elapsed_time(tm[3])}
        flops = 3 * 141.98e9 * per_gpu                 # fwd + dgrad + wgrad, convs only
        achieved_tf = flops / (ms_per_step * 1e-3) / 1e12
        out = {"metric": metric[0], "value": value, "unit": metric[1], "n_gpus": world, "steps": K, "warmup": W,
               "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
               "dtype": "bf16", "data": "synthetic", "config": cfg,
               "roofline": {"bound": "tensor", "kernel": "conv_igemm_kernel (fwd + dgrad) and conv_wgrad_kernel, whole step",
                            "achieved": achieved_tf, "peak": pk["bf16_tflops_sustained"], "unit": "TFLOP/s",
                            "frac": achieved_tf / pk["bf16_tflops_sustained"], "traffic": None,
                            "peak_source": pk["src"] + " (sustained cuBLAS bf16)", "algorithmic_flops": flops,
                            "stage_ms": stage_ms},
               "e2e": {"value": e2e_val, "unit": metric[1], "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                       "ms_per_step": float(te.item())},
               "gpu_launches": int(launches), "clocks": clocks}
        if world == 1 and not args.no_cpu_baseline:
            out["cpu_baseline"], _ = cpu_train(steps=1, batch=2)
        print(json.dumps(out))
        if world > 1:
            dist.destroy_process_group()
        return
    if args.workload == "detect":
        tm = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        run(x, tm)
        torch.cuda.synchronize()
        stage_ms = {"conv_decode": tm[0].elapsed_time(tm[1]), "filter_topk": tm[1].elapsed_time(tm[2]),
                    "rnms": tm[2].elapsed_time(tm[3])}
        flops = 141.98e9 * per_gpu                      # conv MACs*2 per 608x608 image
        achieved_tf = flops / (stage_ms["conv_decode"] * 1e-3) / 1e12
        out = {"metric": metric[0], "value": value, "unit": metric[1], "n_gpus": world, "steps": K, "warmup": W,
               "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
               "dtype": "bf16", "data": "synthetic", "config": cfg,
               "roofline": {"bound": "tensor", "kernel": "conv_igemm_kernel (75 launches/forward, whole conv stack incl. "
                            "first-layer direct conv and decode)", "achieved": achieved_tf,
                            "peak": pk["bf16_tflops_sustained"], "unit": "TFLOP/s",
                            "frac": achieved_tf / pk["bf16_tflops_sustained"], "traffic": None,
                            "peak_source": pk["src"] + " (sustained cuBLAS bf16)", "algorithmic_flops": flops,
                            "stage_ms": stage_ms},
               "e2e": {"value": e2e_val, "unit": metric[1], "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                       "ms_per_step": float(te.item())},
               "gpu_launches": int(launches), "clocks": clocks}
        print(json.dumps(out))
        if world > 1:
            dist.destroy_process_group()
        return
    if args.workload == "riou":
        kern_ms = sorted(per_step)[len(per_step) // 2]   # one launch per step: the step IS the kernel
        kernel = "riou_pairwise_kernel"
    else:
        kern_ms = sorted(per_step)[len(per_step) // 2]
        kernel = "rnms pipeline (rnms_mask_kernel dominant; see profiles/)"
    achieved = alg_bytes / (kern_ms * 1e-3) / 1e9
    traffic = None
    tp = os.path.join(REPO, "profiles", "traffic_%s.json" % args.workload)
    if os.path.exists(tp):
        traffic = json.load(open(tp)).get("dram_bytes_per_launch")
    roof = {"bound": "hbm", "kernel": kernel, "achieved": achieved, "peak": pk["hbm_gbs"], "unit": "GB/s",
            "frac": achieved / pk["hbm_gbs"], "traffic": traffic, "peak_source": pk["src"] + " (burst copy)",
            "algorithmic_bytes": alg_bytes, "kernel_ms": kern_ms}
    if args.workload == "riou":
        # SURVEY.md 8d config 2 also asks for the measured non-zero fraction and the fp32 issue rate: the kernel is bound
        # by instruction issue, not by the 400 MB it writes
        roof["nonzero_fraction"] = float(torch.count_nonzero(outs[0]).item()) / float(n * m)
        if os.path.exists(tp):
            ti = json.load(open(tp)).get("thread_instr_per_launch")
            if ti:
                lane_rate = ti / (kern_ms * 1e-3)
                lane_peak = 148 * 128 * 1.965e9           # fp32 lanes x boost clock
                roof["lane_ops_per_s"] = lane_rate
                roof["lane_ops_frac_of_issue_peak"] = lane_rate / lane_peak
    else:
        kept = pkg.r_nms(d, 0.5)
        roof["kept"] = int(len(kept))
        roof["iou_evals_per_s_upper_triangle"] = 0.5 * n * (n - 1) / (kern_ms * 1e-3)

    out = {"metric": metric[0], "value": value, "unit": metric[1], "n_gpus": world, "steps": K, "warmup": W,
           "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
           "dtype": "f32", "data": "synthetic", "config": cfg, "roofline": roof,
           "e2e": {"value": e2e_val, "unit": metric[1], "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                   "ms_per_step": float(te.item())},
           "gpu_launches": int(launches), "clocks": clocks}
    if world == 1 and not args.no_cpu_baseline:
        if args.workload == "riou":
            out["cpu_baseline"], _ = cpu_riou(sample_rows=10000, steps=2)
        else:
            out["cpu_baseline"], _ = cpu_rnms(20000)
    if world == 1 and args.workload == "riou" and not args.no_also:
        # the other two metrics BASELINE.json names, measured by the same script in child processes AFTER the primary
        # measurement (summary only; run `bench.py --workload rnms|detect` for their full lines)
        also = {}
        for wl, extra in (("rnms", ["--steps", "10", "--warmup", "3"]), ("detect", ["--steps", "3", "--warmup", "3"]),
                          ("train", ["--steps", "3", "--warmup", "3"])):
            try:
                r = subprocess.run([sys.executable, os.path.abspath(__file__), "--workload", wl, "--no-cpu-baseline"] + extra,
                                   stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=600)
                j = json.loads(r.stdout.strip().split("\n")[-1])
                also[wl] = {k: j.get(k) for k in ("metric", "value", "unit", "ms_per_step", "dtype", "gpu_launches")}
                also[wl]["e2e"] = j.get("e2e", {}).get("value")
                also[wl]["roofline_frac"] = j.get("roofline", {}).get("frac")
                also[wl]["roofline_bound"] = j.get("roofline", {}).get("bound")
                if "stage_ms" in j.get("roofline", {}):
                    also[wl]["stage_ms"] = j["roofline"]["stage_ms"]
            except Exception as e:  # never let a secondary workload break the primary line
                also[wl] = {"error": repr(e)[:200]}
        out["other_workloads"] = also
    print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
