"""Who is right where the parity-mode gradients and the fp32 reference golden disagree?  float64 evaluation of the same
graph (stock torch ops on the GPU) as the arbiter: error of OURS vs fp64 and error of the REFERENCE GOLDEN vs fp64."""
import os, sys, copy
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO); sys.path.insert(0, os.path.join(REPO, "tests"))
import numpy as np, torch
import helpers, bench
import rotate_yolov3_b200 as pkg
from rotate_yolov3_b200 import cfgs
g = np.load(os.path.join(helpers.GOLDEN, "darknet_train_golden.npz"), allow_pickle=True)
text = cfgs.yolov3_cfg(width=160, height=128, classes=1, anchors=helpers.SMALL_ANCHORS, n_anchors=6)
def mk(precision):
    m = pkg.Darknet(text, {"context_factor": 1.0}, precision=precision)
    helpers.init_darknet_weights(m, seed=321)
    return m
torch.backends.cudnn.allow_tf32 = False; torch.backends.cuda.matmul.allow_tf32 = False   # stock fp32, not TF32
x = torch.from_numpy(g["x"]).cuda()
gs = [torch.from_numpy(g["g%d" % k]).cuda() for k in range(3)]
# ours
m = mk("parity").cuda().train()
ps = m(x)
(sum((p * gg).sum() for p, gg in zip(ps, gs)) / 100.0).backward()
ours = {n: p.grad.detach().double().reshape(-1) for n, p in m.named_parameters()}
# float64 truth
m64 = mk("bf16").double().cuda().train()
ps64 = bench.torch_port_forward(m64, x.double(), True)
(sum((p * gg.double()).sum() for p, gg in zip(ps64, gs)) / 100.0).backward()
truth = {n: p.grad.detach().reshape(-1) for n, p in m64.named_parameters()}
# fp32 torch on GPU (a second fp32 opinion, different summation order than the CPU reference)
m32 = mk("bf16").cuda().train()
ps32 = bench.torch_port_forward(m32, x, True)
(sum((p * gg).sum() for p, gg in zip(ps32, gs)) / 100.0).backward()
gpu32 = {n: p.grad.detach().double().reshape(-1) for n, p in m32.named_parameters()}
print("forward heads: ours vs fp64 %s | gpu fp32 vs fp64 %s" % (
    ["%.1e" % float((a.double() - b).abs().max() / b.abs().max()) for a, b in zip(ps, ps64)],
    ["%.1e" % float((a.double() - b).abs().max() / b.abs().max()) for a, b in zip(ps32, ps64)]))
rows = []
for name, norm, idx, smp in zip(g["names"], g["norms"], g["sample_idx"], g["samples"]):
    name = str(name)
    t = truth[name]
    sc = float(t.abs().max())
    e_ours = float((ours[name] - t).abs().max()) / sc
    e_gpu32 = float((gpu32[name] - t).abs().max()) / sc
    ii = torch.from_numpy(np.asarray(idx, dtype=np.int64)).cuda()
    e_gold = float((torch.from_numpy(np.asarray(smp, dtype=np.float64)).cuda() - t[ii]).abs().max()) / sc
    rows.append((int(name.split(".")[1]), name, e_ours, e_gold, e_gpu32))
print("%-42s %10s %10s %10s   (max |err| / max |fp64 gradient|; golden on its 256 samples)" % ("parameter", "ours", "cpu-fp32", "gpu-fp32"))
for i, name, a, b, c in rows:
    if name.endswith("Conv2d.weight") or name.endswith("Conv2d.bias"):
        print("%-42s %10.2e %10.2e %10.2e" % (name, a, b, c))
for kind in ("Conv2d.weight", "BatchNorm2d.weight", "BatchNorm2d.bias", "activation.weight"):
    sel = [(a, b, c) for i, n, a, b, c in rows if n.endswith(kind)]
    print("worst %-20s ours %.2e   reference cpu fp32 %.2e   torch gpu fp32 %.2e" % (kind, max(s[0] for s in sel), max(s[1] for s in sel), max(s[2] for s in sel)))

print("---- where are the large deviations of OURS? (elements with |err| > 1e-3 * max)")
for blk in (80, 60, 89, 88, 44, 90, 92):
    name = "module_list.%d.Conv2d.weight" % blk
    t = truth[name]; o = ours[name]
    shp = tuple(dict(m.named_parameters())[name].shape)
    e = (o - t).abs() / float(t.abs().max())
    bad = (e > 1e-3).nonzero().view(-1)
    print(name, shp, "bad elements:", int(bad.numel()), "of", e.numel(), " max err %.3e" % float(e.max()))
    if bad.numel():
        idx = torch.stack(torch.unravel_index(bad, shp), 1).cpu().numpy()
        import collections
        print("   distinct co:", len(set(idx[:, 0])), " distinct ci:", len(set(idx[:, 1])), " taps:", sorted(collections.Counter((int(a), int(b)) for a, b in idx[:, 2:4]).items())[:9])
        print("   first few (co, ci, kh, kw, ours, truth):", [(int(a), int(b), int(c), int(d), float(o[bad[k]]), float(t[bad[k]])) for k, (a, b, c, d) in enumerate(idx[:6])])
        cnt_co = collections.Counter(int(a) for a in idx[:, 0]).most_common(5)
        cnt_ci = collections.Counter(int(a) for a in idx[:, 1]).most_common(5)
        print("   most common co:", cnt_co, " ci:", cnt_ci)
