"""time the BN / PReLU element-wise kernels on one layer-sized tensor (CUDA events, L2-busting rotation of buffers)"""
import sys, os, ctypes, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rotate_yolov3_b200 as pkg
from rotate_yolov3_b200 import layout as L
lib, P = pkg._lib.lib, pkg._lib.ptr
dev = torch.device("cuda")
B, H, W, C = [int(v) for v in sys.argv[1].split(",")]
cs = L.round_up(C, 32)
NB = 3
zs = [torch.randn((B, H + 2, W + 2, cs), device=dev).to(torch.bfloat16) for _ in range(NB)]
dys = [torch.randn((B, H + 2, W + 2, cs), device=dev).to(torch.bfloat16) for _ in range(NB)]
ys = [torch.zeros((B, H + 2, W + 2, cs), dtype=torch.bfloat16, device=dev) for _ in range(NB)]
sums = torch.zeros(2 * C + 1, device=dev)
scale = torch.rand(C, device=dev) + 0.5; shift = torch.randn(C, device=dev); mean = torch.randn(C, device=dev); invstd = torch.rand(C, device=dev) + 0.5
st = pkg._lib.stream_ptr(dev)
nbytes = B * H * W * C * 2
def t(fn, n=9):
    for i in range(3): fn(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(n): fn(i)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
r = {}
r["stats"] = t(lambda i: lib.ryolo_bn_stats(P(zs[i % NB]), cs, B, H, W, C, P(sums), st))
r["fwd"] = t(lambda i: lib.ryolo_bn_act_fwd(P(zs[i % NB]), cs, B, H, W, C, P(scale), P(shift), 0.1, 1, None, 0, P(ys[i % NB]), cs, 0, None, st))
r["bwd"] = t(lambda i: lib.ryolo_bn_act_bwd(P(dys[i % NB]), cs, 0, P(zs[i % NB]), cs, B, H, W, C, P(scale), P(shift), P(mean), P(invstd), 0.1, 1, 1, P(sums), None, 0, 0, None, st))
print("%s items=%s | stats %.0f us (%.2f TB/s) | fwd %.0f us (%.2f TB/s) | bwd reduce+apply %.0f us (%.2f TB/s over 5 passes)" % (
    sys.argv[1], os.environ.get("RYOLO_BN_ITEMS", "16384"), r["stats"], nbytes / r["stats"] / 1e6, r["fwd"], 2 * nbytes / r["fwd"] / 1e6,
    r["bwd"], 5 * nbytes / r["bwd"] / 1e6))
