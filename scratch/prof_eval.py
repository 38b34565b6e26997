import sys, torch
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import helpers, rotate_yolov3_b200 as pkg
from rotate_yolov3_b200 import cfgs
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
m = pkg.Darknet(cfgs.yolov3_cfg(), {'context_factor': 1.0}); helpers.init_darknet_weights(m, 1); m = m.cuda().eval()
x = torch.rand(B, 3, 608, 608, device='cuda')
with torch.no_grad():
    m(x); m(x); torch.cuda.synchronize()
    torch.cuda.profiler.start(); m(x); torch.cuda.synchronize(); torch.cuda.profiler.stop()
