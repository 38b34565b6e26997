import os, sys, numpy as np, torch, torch.nn.functional as F
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
from helpers import SMALL_ANCHORS, init_darknet_weights
import rotate_yolov3_b200 as pkg
from rotate_yolov3_b200 import cfgs
H, W, B = 128, 160, 4
text = cfgs.yolov3_cfg(width=W, height=H, classes=1, anchors=SMALL_ANCHORS, n_anchors=6)
m = pkg.Darknet(text, {"context_factor": 1.0}, arc="default"); init_darknet_weights(m, seed=321); m = m.cuda().train()
g = torch.Generator().manual_seed(11)
x = torch.rand(B, 3, H, W, generator=g).cuda()

def torch_forward(m, x, quant=False):
    q = (lambda t: t.to(torch.bfloat16).float()) if quant else (lambda t: t)
    outs, heads = [], []
    for i, (d, mod) in enumerate(zip(m.module_defs, m.module_list)):
        t = d["type"]
        if t == "convolutional":
            k = mod.Conv2d.weight.shape[-1]
            y = F.conv2d(q(x), q(mod.Conv2d.weight), mod.Conv2d.bias, stride=int(d["stride"]), padding=(k-1)//2)
            if hasattr(mod, "BatchNorm2d"):
                y = F.batch_norm(q(y) if quant else y, None, None, mod.BatchNorm2d.weight, mod.BatchNorm2d.bias, training=True, eps=1e-5)
            if hasattr(mod, "activation"):
                y = F.prelu(y, mod.activation.weight)
            x = y
            if m.module_defs[i+1]["type"] == "yolo": heads.append(y)
        elif t == "shortcut": x = x + outs[i + int(d["from"])]
        elif t == "upsample": x = F.interpolate(x, scale_factor=2, mode="nearest")
        elif t == "route":
            ls = [int(v) for v in d["layers"].split(",")]; ls = [l if l > 0 else i + l for l in ls]
            x = torch.cat([outs[l] for l in ls], 1) if len(ls) > 1 else outs[ls[0]]
        outs.append(x)
    res = []
    for hd, yi in zip(heads, m.yolo_layers):
        layer = m.module_list[yi]
        res.append(hd.view(hd.shape[0], layer.na, layer.nc+6, hd.shape[2], hd.shape[3]).permute(0,1,3,4,2).contiguous())
    return res

ps_t = torch_forward(m, x)
Gs = [torch.randn(p.shape, generator=torch.Generator().manual_seed(5+k)).cuda() for k, p in enumerate(ps_t)]
loss_t = sum((p*gg).sum() for p, gg in zip(ps_t, Gs))/100
loss_t.backward()
ref = {n: p.grad.clone() for n, p in m.named_parameters()}
for p in m.parameters(): p.grad = None
ps_q = torch_forward(m, x, quant=True)
(sum((p*gg).sum() for p, gg in zip(ps_q, Gs))/100).backward()
refq = {n: p.grad.clone() for n, p in m.named_parameters()}
for k,(a,b) in enumerate(zip(ps_q, ps_t)):
    print('torch-bf16-emulation vs fp32: head', k, 'rms rel', float((a-b).pow(2).mean().sqrt()/b.abs().max()))
# restore BN running stats irrelevant; zero grads
for p in m.parameters(): p.grad = None
ps = m(x)
for k,(a,b) in enumerate(zip(ps, ps_q)):
    print('ours vs emulation: head', k, 'rms rel', float((a-b).pow(2).mean().sqrt()/b.abs().max()))
for k,(a,b) in enumerate(zip(ps, ps_t)):
    print('head', k, 'rms rel', float((a-b).pow(2).mean().sqrt()/b.abs().max()), 'cos', float((a*b).sum()/a.norm()/b.norm()))
loss = sum((p*gg).sum() for p, gg in zip(ps, Gs))/100
loss.backward()
for n, p in m.named_parameters():
    if n.endswith('Conv2d.weight') or n.endswith('activation.weight') or n.endswith('BatchNorm2d.bias'):
        a, b = p.grad.float().reshape(-1), ref[n].reshape(-1)
        cos = float((a*b).sum()/(a.norm()*b.norm()+1e-30))
        c = refq[n].reshape(-1); cosq = float((a*c).sum()/(a.norm()*c.norm()+1e-30)); cosqr = float((b*c).sum()/(b.norm()*c.norm()+1e-30))
        if n.endswith('Conv2d.weight'): print(n.split('.')[1], 'W cos(ours,fp32) %.4f cos(ours,emul) %.4f cos(emul,fp32) %.4f' % (cos, cosq, cosqr), end=' | ')
        elif n.endswith('BatchNorm2d.bias'): print('beta cos %.4f' % cos, end=' | ')
        else: print('slope got %.4g ref %.4g' % (float(a), float(b)))
