import os, sys, numpy as np, torch
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
from helpers import GOLDEN, SMALL_ANCHORS, init_darknet_weights
import rotate_yolov3_b200 as pkg
from rotate_yolov3_b200 import cfgs
g = np.load(os.path.join(GOLDEN, "darknet_train_golden.npz"), allow_pickle=True)
text = cfgs.yolov3_cfg(width=160, height=128, classes=1, anchors=SMALL_ANCHORS, n_anchors=6)
m = pkg.Darknet(text, {"context_factor": 1.0}, arc="default"); init_darknet_weights(m, seed=321); m = m.cuda().train()
x = torch.from_numpy(g["x"]).cuda()
ps = m(x)
for k, p in enumerate(ps):
    want = g["p%d" % k]; got = p.detach().cpu().numpy()
    err = np.abs(got - want); scale = np.abs(want).max()
    cos = (got*want).sum()/np.linalg.norm(got)/np.linalg.norm(want)
    print('head', k, 'max', err.max()/scale, 'rms', np.sqrt((err**2).mean())/scale, 'cos', cos)
loss = sum((p * torch.from_numpy(g["g%d" % k]).cuda()).sum() for k, p in enumerate(ps)) / 100.0
print('loss', float(loss), float(g['loss']))
loss.backward()
names = [str(n) for n in g["names"]]; params = dict(m.named_parameters())
rows = []
for name, norm, idx, smp in zip(names, g["norms"], g["sample_idx"], g["samples"]):
    grad = params[name].grad
    if grad is None: print('NO GRAD', name); continue
    got = grad.reshape(-1)[torch.from_numpy(idx.astype(np.int64)).cuda()].float().cpu().numpy(); smp = smp.astype(np.float64)
    cos = float((got*smp).sum()/(np.linalg.norm(got)*np.linalg.norm(smp)+1e-30))
    rows.append((name, cos, float(grad.float().norm())/(float(norm)+1e-30)))
for kind in ('Conv2d.weight','Conv2d.bias','BatchNorm2d.weight','BatchNorm2d.bias','activation.weight'):
    r = [x for x in rows if x[0].endswith(kind)]
    c = np.array([x[1] for x in r]); ra = np.array([x[2] for x in r])
    print(kind, 'n', len(r), 'cos min %.4f med %.4f' % (c.min(), np.median(c)), 'ratio min %.3f med %.3f max %.3f' % (ra.min(), np.median(ra), ra.max()))
    worst = sorted(r, key=lambda t: t[1])[:4]; print('   worst', [(w[0].split('.')[1], round(w[1],3), round(w[2],3)) for w in worst])
print('rm0 err', np.abs(m.module_list[0].BatchNorm2d.running_mean.cpu().numpy()-g['rm0']).max(), 'rv0 err', np.abs(m.module_list[0].BatchNorm2d.running_var.cpu().numpy()-g['rv0']).max())
