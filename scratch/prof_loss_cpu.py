import sys, time, torch, cProfile, pstats
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import helpers, bench, rotate_yolov3_b200 as pkg
from rotate_yolov3_b200 import cfgs
from rotate_yolov3_b200.loss import compute_loss
B = 64; dev = torch.device('cuda')
model = pkg.Darknet(cfgs.yolov3_cfg(), dict(bench.TRAIN_HYP), arc="default"); helpers.init_darknet_weights(model, 1)
model.nc, model.hyp = 1, dict(bench.TRAIN_HYP); model = model.to(dev).train()
opt = torch.optim.SGD(model.parameters(), lr=1e-4, momentum=0.97, nesterov=True)
x = torch.rand(B, 3, 608, 608, device=dev); tg = bench.make_targets(B, 5).to(dev)
def step(prof=None):
    opt.zero_grad(set_to_none=True)
    ps = model(x)
    if prof: prof.enable()
    loss, _ = compute_loss(ps, tg.clone(), model, model.hyp)
    if prof: prof.disable()
    loss.backward(); opt.step()
step(); step(); torch.cuda.synchronize()
pr = cProfile.Profile(); step(pr); torch.cuda.synchronize()
pstats.Stats(pr).sort_stats('tottime').print_stats(12)
