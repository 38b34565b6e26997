# per-layer-class conv timings under the RYOLO_CONV_DEBUG measurement knob (timing experiments; outputs are wrong with knobs on)
import sys, os, torch, ctypes, collections
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import helpers, rotate_yolov3_b200 as pkg
from rotate_yolov3_b200 import cfgs, _lib
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
m = pkg.Darknet(cfgs.yolov3_cfg(), {'context_factor': 1.0}); helpers.init_darknet_weights(m, 1); m = m.cuda().eval()
x = torch.rand(B, 3, 608, 608, device='cuda')
with torch.no_grad():
    m(x); torch.cuda.synchronize()
    plan = m._plan; lib = _lib.lib; stream = _lib.stream_ptr(x.device)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    agg = collections.OrderedDict()
    for kind, a in plan['steps']:
        if kind != 'conv': continue
        d = a['desc']
        def run():
            lib.ryolo_conv_bn_act_fwd(ctypes.byref(d), ctypes.c_void_p(a['x']), _lib.ptr(a['w']), _lib.ptr(a['b']), ctypes.c_void_p(a['r']) if a['r'] else None, ctypes.c_void_p(a['y']), None, 0, stream)
        run(); e0.record()
        for _ in range(3): run()
        e1.record(); torch.cuda.synchronize()
        agg.setdefault((d.cin, d.cout, d.ksize, d.in_h), []).append(e0.elapsed_time(e1) / 3)
    print('DBG', os.environ.get('RYOLO_CONV_DEBUG', '0'), ' '.join('%d>%d.k%d@%d:%.0f' % (*k, 1e3 * sum(v) / len(v)) for k, v in agg.items()), 'total %.2f' % sum(sum(v) for v in agg.values()))
