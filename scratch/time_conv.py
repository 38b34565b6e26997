import sys, time, torch, json
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import helpers, rotate_yolov3_b200 as pkg
from rotate_yolov3_b200 import cfgs
B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
m = pkg.Darknet(cfgs.yolov3_cfg(), {'context_factor': 1.0}); helpers.init_darknet_weights(m, 1); m = m.cuda().eval()
import os
m.use_cuda_graph = os.environ.get('GRAPH','0') == '1'
x = torch.rand(B, 3, 608, 608, device='cuda')
with torch.no_grad():
    io, ps = m(x); torch.cuda.synchronize()
    print('io', tuple(io.shape), 'mem GB', torch.cuda.max_memory_allocated()/2**30, 'finite', bool(torch.isfinite(ps[0]).all()))
    for _ in range(2): m(x)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    N = 5
    for _ in range(N): m(x)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / N
    fl = 141.98e9 * B
    print(json.dumps({'batch': B, 'ms': ms, 'img_s': B / ms * 1e3, 'tflops': fl / ms / 1e9}))
    # per-step timing
    plan = m._plan
    import ctypes
    from rotate_yolov3_b200 import _lib
    lib = _lib.lib; stream = _lib.stream_ptr(x.device)
    rows = []
    for kind, a in plan['steps']:
        if kind != 'conv': continue
        d = a['desc']
        e0.record()
        for _ in range(3):
            lib.ryolo_conv_bn_act_fwd(ctypes.byref(d), ctypes.c_void_p(a['x']), _lib.ptr(a['w']), _lib.ptr(a['b']), ctypes.c_void_p(a['r']) if a['r'] else None, ctypes.c_void_p(a['y']), None, 0, stream)
        e1.record(); torch.cuda.synchronize()
        t = e0.elapsed_time(e1) / 3
        oh = (d.in_h + d.stride - 1)//d.stride; ow = (d.in_w + d.stride - 1)//d.stride
        flops = 2.0 * B * oh * ow * d.cout * d.cin * d.ksize * d.ksize
        rows.append((d.cin, d.cout, d.ksize, d.stride, d.in_h, t, flops / t / 1e9))
    import collections
    agg = collections.OrderedDict()
    for r in rows:
        k = r[:5]; agg.setdefault(k, []).append(r)
    tot = sum(r[5] for r in rows)
    print('conv total ms', tot)
    for k, v in agg.items():
        t = sum(r[5] for r in v)
        print('cin %4d cout %4d k%d s%d h%3d  x%2d  %.3f ms each  %.1f TFLOP/s (useful)  %.1f%% of conv time' % (*k, len(v), t/len(v), v[0][6], 100*t/tot))
