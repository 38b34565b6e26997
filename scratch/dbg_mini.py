import os, sys, numpy as np, torch
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
from helpers import GOLDEN, mini_cfg, init_darknet_weights
import rotate_yolov3_b200 as pkg
g = np.load(os.path.join(GOLDEN, "mini_train_golden.npz"))
m = pkg.Darknet(mini_cfg(64, 48), {"context_factor": 1.0}, arc="default"); init_darknet_weights(m, seed=77); m = m.cuda().train()
x = torch.from_numpy(g["x"]).cuda()
ps = m(x)
loss = sum((p * torch.from_numpy(g["g%d" % k]).cuda()).sum() for k, p in enumerate(ps)) / 10.0
loss.backward()
for name, prm in m.named_parameters():
    want = torch.from_numpy(g["grad:" + name]).cuda().reshape(-1).double(); got = prm.grad.reshape(-1).double()
    cos = float((got * want).sum() / (got.norm() * want.norm() + 1e-30)); ratio = float(got.norm() / (want.norm() + 1e-30))
    print('%-40s cos %.5f ratio %.4f %s' % (name, cos, ratio, ('got %.4g want %.4g' % (float(got), float(want))) if got.numel()==1 else ''))
