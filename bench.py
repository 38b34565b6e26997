#!/usr/bin/env python
"""bench.py -- one JSON line per run (contract in the task statement / DESIGN.md "Measurement").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload train|riou|rnms|detect] [--precision bf16|parity]
                    [--impl ours|reference]

Default = BASELINE.json configs[3], the workload its metric leads with: Darknet-53 (cfg/yolov3.cfg graph, 216 anchors)
training step -- forward (batch-statistics BN) + compute_loss + backward + SGD -- on 64 synthetic 608 x 608 images,
images/s.  Under torchrun the 64 images shard over the ranks (strong scaling, per-replica BN like the reference) and the
bucketed NCCL all-reduce of the 62.4 M fp32 gradients overlaps backward.  The default line also carries, at EVERY N,
`other_workloads`: the same step in parity precision, rotated IoU 10k x 10k (configs[1]), rotated NMS 20k boxes
(configs[2]) and detect e2e (configs[4]).
A "step" is one pass of the hot path over one batch of synthetic input.  `value` is measured with inputs resident in
HBM (CUDA events per step, max over ranks); `e2e` goes through the public Python API with pinned HOST buffers (H2D + D2H
inside the timed region; for the training step ONE timed region around K/2 consecutive steps of the loop a user writes with
`parallel.DevicePrefetcher`: batch i+1 copied under step i, batch 0 exposed, the loss on the host every step).  `roofline`: tensor pipe for the conv workloads (algorithmic FLOPs / measured sustained cuBLAS
bf16 rate), HBM for IoU / NMS (algorithmic bytes / measured copy bandwidth).  `cpu_baseline` / `--impl reference` = the
reference's own code on the host cores where it can be built (oracle/_ref: IoU, NMS), else a port executed by stock
PyTorch CPU kernels (conv workloads; labelled).  The only place bench.py touches oracle/ is those CPU legs and the
reference-CUDA-kernel baseline of the NMS line.
Measurement knobs (scratch/ experiments, never set by default): RYOLO_BENCH_PER_GPU_BATCH (per-rank load of an N-GPU run on
one GPU), RYOLO_BENCH_PROFILE_RANGE=1 (cudaProfilerStart/Stop around the timed region for `ncu --profile-from-start off`)."""
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


def cpu_quota():
    """cgroup CPU quota in cores (None = unlimited): the affinity mask can show 128 CPUs while the container may only
    run a few of them at a time -- the reason the CPU arm differed 6x between two boxes in round 1"""
    try:
        q, p = open("/sys/fs/cgroup/cpu.max").read().split()
        return None if q == "max" else float(q) / float(p)
    except Exception:
        try:
            q = float(open("/sys/fs/cgroup/cpu/cpu.cfs_quota_us").read())
            p = float(open("/sys/fs/cgroup/cpu/cpu.cfs_period_us").read())
            return None if q <= 0 else q / p
        except Exception:
            return None


def cb_finish(cb):
    """per-thread rate and quota next to every CPU number"""
    cb["per_thread"] = cb["value"] / max(1, cb["cores"])
    cb["cpu_quota_cores"] = cpu_quota()
    cb["affinity_cpus"] = cpu_threads()
    return cb


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
    return cb_finish({"value": pairs / min(times) / 1e6, "unit": "Mpairs/s", "cores": cores, "kind": kind,
                      "sample": "%d x 10000 pairs of the 10k x 10k workload; %s" % (sample_rows, what)}), times


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
    return cb_finish({"value": n / min(times), "unit": "boxes/s", "cores": cores, "kind": kind,
                      "sample": "%d boxes (config-3 generator), thr 0.5, K=%d; upper-triangle IoU + reference serial scan"
                                % (n, k)}), times


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
    pg_w = [p for n, p in model.named_parameters() if "Conv2d.weight" in n]
    pg_o = [p for n, p in model.named_parameters() if "Conv2d.weight" not in n]
    opt = torch.optim.SGD([{"params": pg_o}, {"params": pg_w, "weight_decay": 4.569e-4}], lr=1e-4, momentum=0.97, nesterov=True)
    times = []
    for _ in range(steps):
        t0 = time.perf_counter()
        opt.zero_grad(set_to_none=True)
        ps = torch_port_forward(model, x, True)
        for p, yi in zip(ps, model.yolo_layers):
            layer = model.module_list[yi]
            if (layer.nx, layer.ny) != (p.shape[3], p.shape[2]):
                layer.create_grids((608, 608), (p.shape[3], p.shape[2]), "cpu", torch.float32)
        loss, _ = compute_loss(ps, tg.clone(), model, model.hyp)
        loss.backward()
        opt.step()
        times.append(time.perf_counter() - t0)
    return cb_finish({"value": batch / min(times), "unit": "images/s", "cores": torch.get_num_threads(), "kind": "port",
                      "sample": "%d images 608x608 per step: forward + compute_loss + backward + SGD step of THIS REPO's Darknet module "
                                "tree and loss restatement executed by stock PyTorch CPU kernels (not the reference model: "
                                "/root/reference does not exist on the GPU box; same graph, same fp32 nn ops)" % batch}), times


# ---------------------------------------------------------------------------------------------------------------
METRICS = {"riou": ("rotated-IoU Mpairs/sec", "Mpairs/s"), "rnms": ("RNMS boxes/sec", "boxes/s"),
           "detect": ("608x608 images/sec", "images/s"), "train": ("608x608 images/sec", "images/s")}


def config_of(workload, world, precision="bf16"):
    """`config` of BOTH arms (ours / --impl reference): identical keys and values for the same command line"""
    c = {"workload": WORKLOADS[workload]}
    if workload == "train":
        c.update({"global_batch": 64, "per_gpu_batch": max(1, 64 // world), "parallelism": "dp%d" % world,
                  "precision": precision})
    elif workload == "detect":
        c.update({"global_batch": 32, "per_gpu_batch": max(1, 32 // world), "parallelism": "dp%d" % world,
                  "precision": precision})
    elif workload == "riou":
        c.update({"mode": "iou", "parallelism": "row blocks of A per rank (%d), B replicated, no collective" % world})
    else:
        c.update({"parallelism": "one image per rank (%d replicas), no collective" % world})
    return c


class Ctx:
    pass


def timed_steps(ctx, step, K, W, pre_step=None):
    """W untimed + K timed steps; barrier + synchronize on both sides; per-step CUDA events on the launching stream; the
    job's time is the MAX over ranks of each rank's summed step times."""
    import torch
    import torch.distributed as dist
    for i in range(W):
        if pre_step:
            pre_step(i)
        step(i)
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    sampler = ClockSampler(ctx.local_rank)
    if ctx.rank == 0:
        sampler.start()          # BEFORE the barrier: starting it (nvml init, thread) takes tens of ms on rank 0 only, and a
                                 # rank that enters the loop late makes the others wait inside their first timed all-reduce
    torch.cuda.synchronize()
    ctx.barrier()
    l0 = ctx.lib.ryolo_launch_count() + (ctx.replayed() if getattr(ctx, "replayed", None) else 0)
    prof = os.environ.get("RYOLO_BENCH_PROFILE_RANGE") == "1"      # `ncu --profile-from-start off`: the timed region only
    if prof:
        torch.cuda.profiler.start()
    for i in range(K):
        if pre_step:
            pre_step(i)          # e.g. L2 flush: outside the per-step events
        ev[i][0].record()
        step(i)
        ev[i][1].record()
    ctx.barrier()
    if prof:
        torch.cuda.profiler.stop()
    launches = ctx.lib.ryolo_launch_count() + (ctx.replayed() if getattr(ctx, "replayed", None) else 0) - l0
    clocks = sampler.stop() if ctx.rank == 0 else None
    per_step = [s.elapsed_time(e) for s, e in ev]
    t = torch.tensor([sum(per_step)], dtype=torch.float64, device=ctx.dev)
    if ctx.world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ctx.last_per_step = [round(v, 3) for v in per_step]       # rank 0's own steps (the reported time is the max over ranks)
    return float(t.item()) / K, per_step, int(launches), clocks


def timed_e2e(ctx, step_e2e, K, pre_step=None, pipelined=False):
    """end-to-end time per step through the public API with HOST inputs.  Default: each step is timed on its own (copy in,
    compute, result on the host).  pipelined=True (training with parallel.DevicePrefetcher): ONE timed region around Ke
    consecutive steps -- the first step's copy is issued after the start event, step i+1's copy runs under step i, nothing is
    copied that is not consumed inside the region."""
    import torch
    import torch.distributed as dist
    Ke = max(3, K // 2)
    if pipelined:
        for i in range(3):
            step_e2e(i, 3)
        ctx.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(Ke):
            step_e2e(i, Ke)
        e1.record()
        e1.synchronize()
        ctx.barrier()
        te = torch.tensor([e0.elapsed_time(e1) / Ke], dtype=torch.float64, device=ctx.dev)
        if ctx.world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        return float(te.item())
    for i in range(3):
        if pre_step:
            pre_step(i)
        step_e2e(i)
    ctx.barrier()
    tot = 0.0
    for i in range(Ke):
        if pre_step:
            pre_step(i)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        step_e2e(i)          # ends with a stream synchronize (the D2H result is on the host)
        e1.record()
        e1.synchronize()
        tot += e0.elapsed_time(e1)
    ctx.barrier()
    te = torch.tensor([tot / Ke], dtype=torch.float64, device=ctx.dev)
    if ctx.world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    return float(te.item())


def base_line(ctx, wl, K, W, ms_per_step, value, scaling, dtype, cfg, roof, e2e, launches, clocks, notes=None):
    return {"metric": METRICS[wl][0], "value": value, "unit": METRICS[wl][1], "n_gpus": ctx.world, "steps": K, "warmup": W,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": scaling, "vs_baseline": None, "dtype": dtype,
            "data": "synthetic", "config": cfg, "notes": notes or {}, "roofline": roof, "e2e": e2e, "gpu_launches": launches,
            "clocks": clocks}


# ---------------------------------------------------------------------------------------------------------------
def wl_riou(ctx, K, W, full=True):
    import torch
    import helpers
    pkg, dev, pk = ctx.pkg, ctx.dev, ctx.pk
    n = m = 10000
    a_h = helpers.gen_boxes(n, 100 * ctx.rank).pin_memory()   # rank r owns row block r of a (world*10k) x 10k problem
    b_h = helpers.gen_boxes(m, 1).pin_memory()
    a, b = a_h.to(dev), b_h.to(dev)
    outs = [torch.empty((n, m), dtype=torch.float32, device=dev) for _ in range(2)]
    out_h = torch.empty((n, m), dtype=torch.float32).pin_memory()
    alg_bytes = 20 * (n + m) + 4 * n * m      # 400 400 000 B per call

    def step(i):
        pkg.rotated_iou_matrix(a, b, out=outs[i & 1])

    def step_e2e(i):
        ad = a_h.to(dev, non_blocking=True)
        bd = b_h.to(dev, non_blocking=True)
        pkg.rotated_iou_matrix(ad, bd, out=outs[i & 1])
        out_h.copy_(outs[i & 1], non_blocking=True)
        torch.cuda.current_stream().synchronize()
    ms, per_step, launches, clocks = timed_steps(ctx, step, K, W)
    e2e_ms = timed_e2e(ctx, step_e2e, K)
    if ctx.rank != 0:
        return None
    kern_ms = sorted(per_step)[len(per_step) // 2]      # one launch per step: the step IS the kernel
    achieved = alg_bytes / (kern_ms * 1e-3) / 1e9
    traffic, ti = None, None
    tp = os.path.join(REPO, "profiles", "traffic_riou.json")
    if os.path.exists(tp):
        tj = json.load(open(tp))
        traffic, ti = tj.get("dram_bytes_per_launch"), tj.get("thread_instr_per_launch")
    roof = {"bound": "hbm", "kernel": "riou_pairwise_kernel", "achieved": achieved, "peak": pk["hbm_gbs"], "unit": "GB/s",
            "frac": achieved / pk["hbm_gbs"], "traffic": traffic, "peak_source": pk["src"] + " (burst copy)",
            "algorithmic_bytes": alg_bytes, "kernel_ms": kern_ms,
            "nonzero_fraction": float(torch.count_nonzero(outs[0]).item()) / float(n * m),
            "oracle_note": "IoU parity is against a float64 convex-clip restatement; the reference's shapely/GEOS arithmetic "
                           "is an un-vendored dependency and stays unpinned (SURVEY.md 8c)"}
    if ti:
        # SURVEY.md 8d config 2 also asks for the fp32 issue rate: the kernel is bound by instruction issue, not by the
        # 400 MB it writes
        roof["lane_ops_per_s"] = ti / (kern_ms * 1e-3)
        roof["lane_ops_frac_of_issue_peak"] = roof["lane_ops_per_s"] / (148 * 128 * 1.965e9)
    cfg = config_of("riou", ctx.world)
    notes = {"l2": "each step streams a 400 MB output (> 126 MB L2) into alternating buffers"}
    out = base_line(ctx, "riou", K, W, ms, n * m * ctx.world / (ms * 1e-3) * 1e-6, "weak", "f32", cfg, roof,
                    {"value": n * m * ctx.world / (e2e_ms * 1e-3) * 1e-6, "unit": "Mpairs/s", "h2d_bytes_per_step": (n + m) * 20,
                     "d2h_bytes_per_step": n * m * 4, "ms_per_step": e2e_ms}, launches, clocks, notes)
    if full and ctx.world == 1 and not ctx.args.no_cpu_baseline:
        out["cpu_baseline"], _ = cpu_riou(sample_rows=10000, steps=2)
    return out


def wl_rnms(ctx, K, W, full=True):
    import numpy as np
    import torch
    import helpers
    pkg, dev, pk = ctx.pkg, ctx.dev, ctx.pk
    n = 20000
    d_h = helpers.gen_dets(n, 2 + 100 * ctx.rank).pin_memory()
    d = d_h.to(dev)
    cbk = (n + 63) // 64
    alg_bytes = 24 * n + 8 * n * cbk * 2       # boxes in + mask write + mask read (reference algorithm)
    flush = torch.empty(192 << 20, dtype=torch.uint8, device=dev)
    keep_h = torch.empty(n, dtype=torch.long).pin_memory()

    def pre(i):
        flush.fill_(i & 0xFF)      # L2 flush between timed iterations (both the device-resident and the e2e loop)

    def step(i):
        pkg.r_nms(d, 0.5)

    def step_e2e(i):
        dd = d_h.to(dev, non_blocking=True)
        k = pkg.r_nms(dd, 0.5)
        keep_h[:len(k)].copy_(k)
    ms, per_step, launches, clocks = timed_steps(ctx, step, K, W, pre)
    e2e_ms = timed_e2e(ctx, step_e2e, K, pre)
    if ctx.rank != 0:
        return None
    kern_ms = sorted(per_step)[len(per_step) // 2]
    achieved = alg_bytes / (kern_ms * 1e-3) / 1e9
    traffic = None
    tp = os.path.join(REPO, "profiles", "traffic_rnms.json")
    if os.path.exists(tp):
        traffic = json.load(open(tp)).get("dram_bytes_per_launch")
    kept = pkg.r_nms(d, 0.5)
    roof = {"bound": "hbm", "kernel": "rnms pipeline (rnms_mask_kernel dominant; see profiles/)", "achieved": achieved,
            "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": achieved / pk["hbm_gbs"], "traffic": traffic,
            "peak_source": pk["src"] + " (burst copy)", "algorithmic_bytes": alg_bytes, "kernel_ms": kern_ms,
            "kept": int(len(kept)), "iou_evals_per_s_upper_triangle": 0.5 * n * (n - 1) / (kern_ms * 1e-3),
            "note": "HBM is the roofline north_star names; the binding resources are fp32 issue in the exact IoU stage and "
                    "the serial greedy scan (DESIGN.md section 3)"}
    # second baseline: the REFERENCE CUDA kernel (rotate_polygon_nms_kernel.cu:262-308 compiled by nvcc for this GPU,
    # oracle/_ref) on the same boxes -- its mask launch alone, device-resident, and its whole nms_cuda() from host memory
    ref = helpers.ref_lib("cuda")
    if ref is not None:
        try:
            order = torch.argsort(d[:, 5], descending=True)
            sb = d[order].contiguous()
            mask = torch.empty((n, cbk), dtype=torch.int64, device=dev)
            ts = []
            for i in range(4):
                flush.fill_(i)
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                rc = ref.ref_cuda_mask(ctypes.c_void_p(sb.data_ptr()), n, ctypes.c_float(0.5), ctypes.c_void_p(mask.data_ptr()))
                ts.append(time.perf_counter() - t0)
                assert rc == 0
            dn = np.ascontiguousarray(d_h.numpy())
            keep = np.empty(n, np.int64)
            tw = []
            for i in range(3):
                t0 = time.perf_counter()
                kk = ref.ref_cuda_rnms(helpers.P(dn), n, ctypes.c_float(0.5), keep.ctypes.data_as(helpers.I64P))
                tw.append(time.perf_counter() - t0)
            same = bool(kk == len(kept) and np.array_equal(keep[:kk], kept.cpu().numpy()))
            roof["reference_cuda_kernel"] = {
                "mask_kernel_ms": 1e3 * min(ts[1:]), "whole_nms_cuda_ms": 1e3 * min(tw[1:]),
                "boxes_per_s_whole": n / min(tw[1:]), "keep_list_identical_to_ours": same,
                "what": "the reference's rotate_nms_kernel on this B200 (all N^2 pairs) / its nms_cuda(): sort, kernel, 50 MB "
                        "blocking D2H, host scan (oracle/_ref/libref_rnms_cuda.so)"}
        except Exception as e:   # the baseline must never break the line
            roof["reference_cuda_kernel"] = {"error": repr(e)[:200]}
    cfg = config_of("rnms", ctx.world)
    notes = {"l2": "50 MB mask rewritten every step; 192 MB flush buffer written between timed steps (device and e2e loops)"}
    out = base_line(ctx, "rnms", K, W, ms, n * ctx.world / (ms * 1e-3), "weak", "f32", cfg, roof,
                    {"value": n * ctx.world / (e2e_ms * 1e-3), "unit": "boxes/s", "h2d_bytes_per_step": n * 24,
                     "d2h_bytes_per_step": 8 * int(len(kept)), "ms_per_step": e2e_ms}, launches, clocks, notes)
    if full and ctx.world == 1 and not ctx.args.no_cpu_baseline:
        out["cpu_baseline"], _ = cpu_rnms(20000)
    return out


def wl_train(ctx, K, W, full=True, precision="bf16"):
    """BASELINE configs[3]: Darknet-53 training step, global batch 64 x 608 x 608 synthetic, compute_loss of the reference
    (no rotated IoU in it), SGD nesterov; images shard across ranks, per-replica BN."""
    import torch
    import helpers
    from rotate_yolov3_b200 import cfgs, parallel
    from rotate_yolov3_b200.loss import compute_loss
    pkg, dev, pk, world = ctx.pkg, ctx.dev, ctx.pk, ctx.world
    per_gpu = max(1, 64 // world)
    if os.environ.get("RYOLO_BENCH_PER_GPU_BATCH"):        # experiment knob (scratch/): the per-rank load of an N-GPU run on one GPU
        per_gpu = int(os.environ["RYOLO_BENCH_PER_GPU_BATCH"])
    net = pkg.Darknet(cfgs.yolov3_cfg(), dict(TRAIN_HYP), arc="default", precision=precision)
    helpers.init_darknet_weights(net, seed=1)
    net.nc, net.hyp = 1, dict(TRAIN_HYP)
    net = net.to(dev).train()
    net.use_cuda_graph = (precision == "bf16") and not ctx.args.no_graph
    model = parallel.DistributedDataParallel(net) if world > 1 else net     # train.py:175
    pg_w = [p for n, p in net.named_parameters() if "Conv2d.weight" in n]
    pg_o = [p for n, p in net.named_parameters() if "Conv2d.weight" not in n]
    groups = [{"params": pg_o}, {"params": pg_w, "weight_decay": 4.569e-4}]      # train.py:70-82 param groups, cfg/hyp_template.py
    try:        # the framework's single-kernel multi-tensor SGD (same update rule as train.py's optim.SGD)
        opt = torch.optim.SGD(groups, lr=1e-4, momentum=0.97, nesterov=True, fused=True)
    except (TypeError, RuntimeError):
        opt = torch.optim.SGD(groups, lr=1e-4, momentum=0.97, nesterov=True)
    x_h = torch.rand(per_gpu, 3, 608, 608, generator=torch.Generator().manual_seed(ctx.rank)).pin_memory()
    tg_h = make_targets(per_gpu, 100 + ctx.rank).pin_memory()
    x, tg = x_h.to(dev), tg_h.to(dev)
    loss_h = torch.empty(1).pin_memory()

    def train_step(xd, td, timers=None):
        opt.zero_grad(set_to_none=True)
        if timers:
            timers[0].record()
        ps = model(xd)
        if timers:
            timers[1].record()
        loss, items = compute_loss(ps, td.clone(), net, net.hyp)
        loss.backward()          # the bucketed all-reduce runs inside, overlapped with the backward kernels
        if timers:
            timers[2].record()
        opt.step()
        if timers:
            timers[3].record()
        return loss

    def step(i):
        train_step(x, tg)

    pf = parallel.DevicePrefetcher(dev)

    def step_e2e(i, n):
        # the loop a user writes with the package's prefetcher: batch i+1 is copied from pinned host memory while step i
        # computes; every step ends with its loss on the host
        if i == 0:
            pf.put(x_h, tg_h)
        xd, td = pf.get()
        if i + 1 < n:
            pf.put(x_h, tg_h)
        loss = train_step(xd, td)
        loss_h.copy_(loss.detach(), non_blocking=True)
        torch.cuda.current_stream().synchronize()
    if precision == "parity":
        K, W = min(K, 3), min(W, 3)
    graph_note = None
    if net.use_cuda_graph:
        # two steps outside the timed region: eager + capture, then the first replay.  If capture is not possible on this
        # software stack, say so and measure the eager path instead of failing the whole line.
        try:
            step(0)
            step(1)
            torch.cuda.synchronize()
        except Exception as e:
            graph_note = "CUDA-graph capture failed (%s); eager launches measured" % repr(e)[:160]
            net.use_cuda_graph = False
            net._tplan = None
            torch.cuda.synchronize()
    # kernels executed through CUDA-graph replays are counted from the capture (the library's launch counter only sees eager calls)
    ctx.replayed = lambda: (net._tplan.replayed_kernels if getattr(net, "_tplan", None) is not None else 0)
    ms, per_step, launches, clocks = timed_steps(ctx, step, K, W)
    ctx.replayed = None
    e2e_ms = timed_e2e(ctx, step_e2e, K, pipelined=True)
    # one more step with stage timers on EVERY rank (it contains the collective)
    tm = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    plan = net._pplan if precision == "parity" else net._tplan
    if hasattr(plan, "time_comm"):
        plan.time_comm = True
    torch.cuda.synchronize()
    h0 = time.perf_counter()
    train_step(x, tg, tm)
    ctx.host_ms_per_step = (time.perf_counter() - h0) * 1e3     # the HOST's time to enqueue one step on an idle device
    torch.cuda.synchronize()
    exposed = getattr(plan, "last_comm_wait_ms", lambda: None)()
    ctx.barrier()
    out = None
    if ctx.rank == 0:
        stage_ms = {"forward": tm[0].elapsed_time(tm[1]), "loss_backward": tm[1].elapsed_time(tm[2]),
                    "sgd": tm[2].elapsed_time(tm[3]), "allreduce_exposed": exposed,
                    "allreduce_sgd": (exposed or 0.0) + tm[2].elapsed_time(tm[3])}
        flops = 3 * 141.98e9 * per_gpu * world            # fwd + dgrad + wgrad, convs only, whole job
        achieved_tf = flops / (ms * 1e-3) / 1e12
        peak = pk["bf16_tflops_sustained"] * world
        cfg = config_of("train", world, precision)
        notes = {}
        notes.update({"collective": "bucketed NCCL all-reduce of 62.4M fp32 gradients (32 MB buckets in backward order), "
                                  "overlapped with backward" if world > 1 else "none (1 rank)",
                    "cuda_graph": bool(net.use_cuda_graph) if graph_note is None else graph_note,
                    "rank0_per_step_ms": getattr(ctx, "last_per_step", None),
                    "rank0_host_enqueue_ms_per_step": round(getattr(ctx, "host_ms_per_step", 0.0), 3),
                    "reserved_sms": (net._ddp or {}).get("reserved_sms") if getattr(net, "_ddp", None) else 0,
                    "arithmetic": "bf16 operands/activations, fp32 accumulate and parameter gradients" if precision == "bf16"
                    else "fp32-grade: 3 bf16 planes per operand, 6 exact-product terms, fp32 round-to-nearest sums, fp64 "
                         "reductions (parity_path.py); matches the reference's fp32 modules to 1e-4",
                    "l2": "activations of one step (tens of GB) exceed the 126 MB L2"})
        roof = {"bound": "tensor", "kernel": "conv_igemm_kernel (fwd + dgrad) and conv_wgrad_kernel, whole step" if
                precision == "bf16" else "conv_px_kernel (fwd + dgrad) and conv_wgrad_kernel over 3 planes, whole step",
                "achieved": achieved_tf, "peak": peak, "unit": "TFLOP/s", "frac": achieved_tf / peak, "traffic": None,
                "peak_source": pk["src"] + " (sustained cuBLAS bf16, x n_gpus)", "algorithmic_flops": flops,
                "stage_ms": stage_ms}
        out = base_line(ctx, "train", K, W, ms, per_gpu * world / (ms * 1e-3), "strong", "bf16" if precision == "bf16" else "f32",
                        cfg, roof, {"value": per_gpu * world / (e2e_ms * 1e-3), "unit": "images/s",
                                    "h2d_bytes_per_step": per_gpu * 3 * 608 * 608 * 4 + tg_h.numel() * 4,
                                    "d2h_bytes_per_step": 4, "ms_per_step": e2e_ms,
                                    "how": "one timed region around K/2 consecutive steps; inputs in pinned host memory, copied by "
                                           "parallel.DevicePrefetcher (batch i+1 under step i, batch 0 exposed); the loss is "
                                           "read to the host every step"}, launches, clocks, notes)
    del model, net, opt, plan
    torch.cuda.empty_cache()
    if out is not None and full and world == 1 and not ctx.args.no_cpu_baseline:
        out["cpu_baseline"], _ = cpu_train(steps=1, batch=2)
    return out


def wl_detect(ctx, K, W, full=True):
    """BASELINE configs[4] shape: eval forward (conv stacks + decode) -> conf filter -> per-image top-20000 -> RNMS.
    Per-GPU batch = 32 / world (images shard across ranks, no collective); random-init Darknet-53 with randomised BN
    statistics / PReLU slopes."""
    import torch
    import helpers
    from rotate_yolov3_b200 import cfgs
    from rotate_yolov3_b200.nms import detect_postprocess
    pkg, dev, pk, world = ctx.pkg, ctx.dev, ctx.pk, ctx.world
    per_gpu = max(1, 32 // world)
    model = pkg.Darknet(cfgs.yolov3_cfg(), {"context_factor": 1.0}, arc="default")
    helpers.init_darknet_weights(model, seed=1)
    model = model.to(dev).eval()
    model.use_cuda_graph = True      # the ~80 launches of one forward replayed as one CUDA graph
    x_h = torch.rand(per_gpu, 3, 608, 608, generator=torch.Generator().manual_seed(ctx.rank)).pin_memory()
    x = x_h.to(dev)
    CAP, CONF, THR = 20000, 0.5, 0.5

    def run(xd, timers=None):
        with torch.no_grad():
            if timers:
                timers[0].record()
            io, _ = model(xd)
            if timers:
                timers[1].record()
            res = detect_postprocess(io, CONF, THR, CAP, timers[2] if timers else None)
            if timers:
                timers[3].record()
        return res

    def step(i):
        run(x)

    counts_h = torch.empty(per_gpu, dtype=torch.int32).pin_memory()

    def step_e2e(i):
        xd = x_h.to(dev, non_blocking=True)
        res = run(xd)
        counts_h.copy_(res["num_keep"], non_blocking=True)
        torch.cuda.current_stream().synchronize()
    ctx.replayed = lambda: model.replayed_kernels
    ms, per_step, launches, clocks = timed_steps(ctx, step, K, W)
    ctx.replayed = None
    e2e_ms = timed_e2e(ctx, step_e2e, K)
    out = None
    if ctx.rank == 0:
        tm = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        run(x, tm)
        torch.cuda.synchronize()
        stage_ms = {"conv_decode": tm[0].elapsed_time(tm[1]), "filter_topk": tm[1].elapsed_time(tm[2]),
                    "rnms": tm[2].elapsed_time(tm[3])}
        flops = 141.98e9 * per_gpu                      # conv MACs*2 per 608x608 image
        achieved_tf = flops / (stage_ms["conv_decode"] * 1e-3) / 1e12
        cfg = config_of("detect", world)
        notes = {"l2": "activations of one forward (~8 GB at batch 32) exceed the 126 MB L2"}
        roof = {"bound": "tensor", "kernel": "conv_igemm_kernel (75 launches/forward, whole conv stack incl. first-layer "
                "conv and decode)", "achieved": achieved_tf, "peak": pk["bf16_tflops_sustained"], "unit": "TFLOP/s",
                "frac": achieved_tf / pk["bf16_tflops_sustained"], "traffic": None,
                "peak_source": pk["src"] + " (sustained cuBLAS bf16)", "algorithmic_flops": flops, "stage_ms": stage_ms}
        out = base_line(ctx, "detect", K, W, ms, per_gpu * world / (ms * 1e-3), "strong", "bf16", cfg, roof,
                        {"value": per_gpu * world / (e2e_ms * 1e-3), "unit": "images/s",
                         "h2d_bytes_per_step": per_gpu * 3 * 608 * 608 * 4, "d2h_bytes_per_step": per_gpu * 4,
                         "ms_per_step": e2e_ms}, launches, clocks, notes)
    del model
    torch.cuda.empty_cache()
    return out


def summarise(j):
    if j is None:
        return None
    s = {k: j.get(k) for k in ("metric", "value", "unit", "ms_per_step", "dtype", "gpu_launches", "steps", "warmup")}
    s["e2e"] = j.get("e2e", {}).get("value")
    r = j.get("roofline", {})
    s["roofline_frac"], s["roofline_bound"] = r.get("frac"), r.get("bound")
    for k in ("stage_ms", "reference_cuda_kernel", "kernel_ms"):
        if k in r:
            s[k] = r[k]
    s["precision"] = j.get("config", {}).get("precision")
    return s


def reference_arm(args, rank):
    """--impl reference: the reference's own CPU implementation of the path on the host cores, rank 0 only"""
    if rank != 0:
        return
    K, W = args.steps, args.warmup
    wl = args.workload
    metric = METRICS[wl]
    if wl == "riou":
        n_run = max(1, min(K, 3))
        cb, times = cpu_riou(sample_rows=10000, steps=n_run + min(W, 1))
        sample = "the full 1e8 pairs per step"
    elif wl == "rnms":
        n_run = max(1, min(K, 3))
        cb, times = cpu_rnms(20000, steps=n_run + min(W, 1))
        sample = "the full 20000 boxes per step"
    elif wl == "train":
        n_run = max(1, min(K, 4))
        cb, times = cpu_train(steps=n_run + min(W, 1), batch=2)
        sample = "bounded sample of 2 of the 64 images per step"
    else:
        print(json.dumps({"impl": "reference", "unavailable": "detect workload has no CPU reference arm (conv port + host "
                          "NMS would take minutes per image); its stages are covered by the train / rnms arms"}))
        return
    warm = min(W, 1)
    timed = times[warm:] if len(times) > warm else times
    ms = 1e3 * sum(timed) / len(timed)                 # mean over the steps actually timed: steps * ms_per_step = wall time
    scale = {"riou": 1e8 / 1e6, "rnms": 20000.0, "train": 2.0}[wl]
    value = scale / (ms * 1e-3)
    cb["value"] = value
    cb["per_thread"] = value / max(1, cb["cores"])
    cb["sample"] = sample + "; " + cb["sample"]
    print(json.dumps({"impl": "reference", "metric": metric[0], "value": value, "unit": metric[1], "n_gpus": args.gpus,
                      "steps": len(timed), "warmup": warm, "ms_per_step": ms, "higher_is_better": True,
                      "scaling": "strong" if wl in ("train", "detect") else "weak", "vs_baseline": None,
                      "dtype": "f32", "data": "synthetic", "config": config_of(wl, args.gpus, args.precision),
                      "notes": {"arithmetic": "fp32 on the host cores; config.precision names the GPU arm's mode"},
                      "cpu_baseline": cb, "steps_requested": K,
                      "e2e": {"value": value, "unit": metric[1], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default="train", choices=["riou", "rnms", "detect", "train"])
    ap.add_argument("--precision", default="bf16", choices=["bf16", "parity"])
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-also", action="store_true", help="skip the secondary workloads of the default run")
    ap.add_argument("--no-graph", action="store_true", help="training step without CUDA-graph replay")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        reference_arm(args, rank)
        return

    import torch
    import torch.distributed as dist
    import rotate_yolov3_b200 as pkg
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local_rank)
    ctx = Ctx()
    ctx.args, ctx.rank, ctx.world, ctx.local_rank = args, rank, world, local_rank
    ctx.dev = torch.device("cuda", local_rank)
    if world > 1:
        # the gradient all-reduce runs NEXT to the backward GEMMs: cap the SMs NCCL may take (at 2 GPUs 8 and 16 channels
        # measure the same, profiles/r02_n2_experiments.txt)
        os.environ.setdefault("NCCL_MAX_NCHANNELS", "16")
        dist.init_process_group("nccl", device_id=ctx.dev)
    ctx.pk, ctx.pkg, ctx.lib = peaks(), pkg, pkg._lib.lib

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
    ctx.barrier = barrier
    K, W = args.steps, max(args.warmup, 3)
    wl = args.workload
    if wl == "train":
        out = wl_train(ctx, K, W, True, args.precision)
    else:
        out = {"riou": wl_riou, "rnms": wl_rnms, "detect": wl_detect}[wl](ctx, K, W, True)
    if wl == "train" and args.precision == "bf16" and not args.no_also:
        # the other workloads BASELINE.json names + the same step in parity precision, on every rank count
        also = {}
        for name, fn in (("train_parity", lambda: wl_train(ctx, 2, 3, False, "parity")),
                         ("riou", lambda: wl_riou(ctx, 10, 3, False)), ("rnms", lambda: wl_rnms(ctx, 10, 3, False)),
                         ("detect", lambda: wl_detect(ctx, 3, 3, False))):
            try:
                r = fn()
                if rank == 0:
                    also[name] = summarise(r)
            except Exception as e:      # a secondary workload must not break the primary line
                torch.cuda.empty_cache()
                if rank == 0:
                    also[name] = {"error": repr(e)[:300]}
        if rank == 0:
            out["other_workloads"] = also
    if rank == 0:
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
