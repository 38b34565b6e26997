# run ONE conv layer of the batch-32 eval plan a few times (ncu target): args cin,cout,k,h [reps]
import sys, torch, ctypes
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import helpers, rotate_yolov3_b200 as pkg
from rotate_yolov3_b200 import cfgs, _lib
key = tuple(int(v) for v in sys.argv[1].split(',')); reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
m = pkg.Darknet(cfgs.yolov3_cfg(), {'context_factor': 1.0}); helpers.init_darknet_weights(m, 1); m = m.cuda().eval()
x = torch.rand(32, 3, 608, 608, device='cuda')
with torch.no_grad():
    m(x); torch.cuda.synchronize()
    lib = _lib.lib; stream = _lib.stream_ptr(x.device)
    for kind, a in m._plan['steps']:
        if kind != 'conv': continue
        d = a['desc']
        if (d.cin, d.cout, d.ksize, d.in_h) != key: continue
        torch.cuda.profiler.start()
        for _ in range(reps):
            lib.ryolo_conv_bn_act_fwd(ctypes.byref(d), ctypes.c_void_p(a['x']), _lib.ptr(a['w']), _lib.ptr(a['b']), ctypes.c_void_p(a['r']) if a['r'] else None, ctypes.c_void_p(a['y']), None, 0, stream)
        torch.cuda.synchronize(); torch.cuda.profiler.stop()
        break
