import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
import helpers
import rotate_yolov3_b200 as pkg
from rotate_yolov3_b200 import cfgs
g = np.load(os.path.join(helpers.GOLDEN, "darknet_train_golden.npz"), allow_pickle=True)
m = pkg.Darknet(cfgs.yolov3_cfg(width=160, height=128, classes=1, anchors=helpers.SMALL_ANCHORS, n_anchors=6), {"context_factor": 1.0}, precision="parity")
helpers.init_darknet_weights(m, seed=321)
m = m.cuda().train()
ps = m(torch.from_numpy(g["x"]).cuda())
loss = sum((p * torch.from_numpy(g["g%d" % k]).cuda()).sum() for k, p in enumerate(ps)) / 100.0
loss.backward()
grads = dict(m.named_parameters())
for name, norm, idx, smp in zip(g["names"], g["norms"], g["sample_idx"], g["samples"]):
    name = str(name)
    i = int(name.split(".")[1])
    gr = grads[name].grad.reshape(-1)
    got = gr[torch.from_numpy(np.asarray(idx, dtype=np.int64)).cuda()].cpu().numpy()
    scale = max(float(np.abs(smp).max()), 1e-30)
    es = float(np.abs(got - smp).max()) / scale
    en = abs(float(gr.norm()) - float(norm)) / float(norm)
    if i >= 55 or i < 6:
        print("%-40s numel %8d  sample err %.2e  norm err %.2e  |ref| %.3e" % (name, gr.numel(), es, en, float(norm)))
