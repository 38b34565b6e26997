"""Experiment (round 2): how accurate is the fp32 accumulation of tcgen05.mma.kind::f16 over long K chains?
Decides the design of the parity-precision conv mode (bf16x3 split needs the accumulator itself to be ~fp32-exact).
bf16-representable operands -> every product is exact in fp32, so the only error is the accumulation."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
from rotate_yolov3_b200 import layout as L

dev = torch.device("cuda")


def run(cin, cout, k, h=19, w=19, batch=2, seed=0, positive=False):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(batch, cin, h, w, generator=g)
    wt = torch.randn(cout, cin, k, k, generator=g) / (cin * k * k) ** 0.5
    if positive:
        x, wt = x.abs(), wt.abs()
    x = x.to(dev).to(torch.bfloat16).float()
    wt = wt.to(dev).to(torch.bfloat16).float()
    bias = torch.zeros(cout, device=dev)
    xb = L.to_padded_nhwc(x, L.round_up(cin, 64))
    desc = L.make_desc(batch, h, w, cin, L.round_up(cin, 64), cout, 0, k, 1, False, 0.0, False, 0, False, True)
    pw = L.pack_weights(desc, wt)
    pb = L.padded_bias(desc, bias)
    y = torch.empty((batch, cout, h, w), device=dev)
    L.conv_fwd(desc, xb.data_ptr(), pw, pb, y.data_ptr(), None, dev)
    torch.cuda.synchronize()
    want = F.conv2d(x.double(), wt.double(), padding=(k - 1) // 2)
    want32 = F.conv2d(x, wt, padding=(k - 1) // 2)
    e = (y.double() - want)
    e32 = (want32.double() - want)
    rms = float(want.pow(2).mean().sqrt())
    print("cin=%d k=%d K=%d pos=%d | tcgen05: max %.3e rms %.3e mean(signed*sign(want)) %.3e | cudnn fp32: max %.3e rms %.3e  (all / out rms %.3f)"
          % (cin, k, cin * k * k, positive, float(e.abs().max()) / rms, float(e.pow(2).mean().sqrt()) / rms,
             float((e * want.sign()).mean()) / rms, float(e32.abs().max()) / rms, float(e32.pow(2).mean().sqrt()) / rms, rms))


torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
for cin, k in ((64, 1), (256, 1), (1024, 1), (128, 3), (512, 3), (1024, 3)):
    run(cin, 128, k)
run(1024, 128, 3, positive=True)
run(256, 128, 3, positive=True)

# emulation of the split: fp32 operands a, w -> (hi, lo) bf16 pairs; a*w ~ ah*wh + ah*wl + al*wh in exact arithmetic
g = torch.Generator().manual_seed(3)
a = torch.randn(4096, 2304, generator=g).to(dev)
b = (torch.randn(2304, 256, generator=g) / 48).to(dev)
ah = a.to(torch.bfloat16).float(); al = (a - ah).to(torch.bfloat16).float()
bh = b.to(torch.bfloat16).float(); bl = (b - bh).to(torch.bfloat16).float()
want = a.double() @ b.double()
rms = float(want.pow(2).mean().sqrt())
for name, got in (("hh", ah.double() @ bh.double()),
                  ("hh+hl+lh", ah.double() @ bh.double() + ah.double() @ bl.double() + al.double() @ bh.double()),
                  ("all4", (ah + al).double() @ (bh + bl).double()),
                  ("fp32 matmul", (a @ b).double())):
    e = got - want
    print("split emulation %-12s max %.3e rms %.3e (rel. to out rms)" % (name, float(e.abs().max()) / rms, float(e.pow(2).mean().sqrt()) / rms))
