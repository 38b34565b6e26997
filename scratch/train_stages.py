import sys, time, torch
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import helpers, bench, rotate_yolov3_b200 as pkg
from rotate_yolov3_b200 import cfgs, train_path
from rotate_yolov3_b200.loss import compute_loss
B = 64; dev = torch.device('cuda')
model = pkg.Darknet(cfgs.yolov3_cfg(), dict(bench.TRAIN_HYP), arc="default"); helpers.init_darknet_weights(model, 1)
model.nc, model.hyp = 1, dict(bench.TRAIN_HYP); model = model.to(dev).train()
opt = torch.optim.SGD(model.parameters(), lr=1e-4, momentum=0.97, nesterov=True)
x = torch.rand(B, 3, 608, 608, device=dev); tg = bench.make_targets(B, 5).to(dev)
ev = {}
def mark(n):
    e = torch.cuda.Event(enable_timing=True); e.record(); ev[n] = (e, time.perf_counter())
orig_bwd = train_path.TrainPlan.backward
def bwd(self, grads):
    mark('pb0'); r = orig_bwd(self, grads); mark('pb1'); return r
train_path.TrainPlan.backward = bwd
def step():
    mark('s0'); opt.zero_grad(set_to_none=True)
    ps = model(x); mark('f1')
    loss, _ = compute_loss(ps, tg.clone(), model, model.hyp); mark('l1')
    loss.backward(); mark('b1'); opt.step(); mark('o1')
for _ in range(3): step()
torch.cuda.synchronize(); t0 = time.perf_counter(); step(); torch.cuda.synchronize(); t1 = time.perf_counter()
names = ['s0','f1','l1','pb0','pb1','b1','o1']
for a, b in zip(names[:-1], names[1:]):
    print('%s->%s gpu %.2f ms  cpu %.2f ms' % (a, b, ev[a][0].elapsed_time(ev[b][0]), 1e3*(ev[b][1]-ev[a][1])))
print('step wall %.2f ms' % (1e3*(t1-t0)))
