import sys, time, torch, warnings, traceback
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import helpers, bench, rotate_yolov3_b200 as pkg
from rotate_yolov3_b200 import cfgs
from rotate_yolov3_b200.loss import compute_loss
B = 8; dev = torch.device('cuda')
model = pkg.Darknet(cfgs.yolov3_cfg(), dict(bench.TRAIN_HYP), arc="default"); helpers.init_darknet_weights(model, 1)
model.nc, model.hyp = 1, dict(bench.TRAIN_HYP); model = model.to(dev).train()
opt = torch.optim.SGD(model.parameters(), lr=1e-4, momentum=0.97, nesterov=True)
x = torch.rand(B, 3, 608, 608, device=dev); tg = bench.make_targets(B, 5).to(dev)
def step():
    opt.zero_grad(set_to_none=True)
    ps = model(x); loss, _ = compute_loss(ps, tg.clone(), model, model.hyp); loss.backward(); opt.step()
step(); step(); torch.cuda.synchronize()
def showwarning(message, category, filename, lineno, file=None, line=None):
    st = [f for f in traceback.extract_stack() if '/root/repo' in f.filename and 'find_syncs' not in f.filename]
    print('SYNC:', str(message)[:80], '|', ' <- '.join('%s:%d' % (f.filename.split('/')[-1], f.lineno) for f in st[-3:]))
warnings.showwarning = showwarning
warnings.simplefilter('always')
torch.cuda.set_sync_debug_mode("warn")
step()
torch.cuda.set_sync_debug_mode("default")
