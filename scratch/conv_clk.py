# does the SM clock under a sustained conv loop depend on the epilogue activity? (power-capped B200)
import sys, os, time, threading, torch, ctypes
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import helpers, rotate_yolov3_b200 as pkg, pynvml
from rotate_yolov3_b200 import cfgs, _lib
pynvml.nvmlInit(); h = pynvml.nvmlDeviceGetHandleByIndex(0)
B = 32
m = pkg.Darknet(cfgs.yolov3_cfg(), {'context_factor': 1.0}); helpers.init_darknet_weights(m, 1); m = m.cuda().eval()
x = torch.rand(B, 3, 608, 608, device='cuda')
want = [tuple(int(v) for v in a.split(',')) for a in sys.argv[1:]] or [(128, 256, 3, 76)]
with torch.no_grad():
    m(x); torch.cuda.synchronize()
    lib = _lib.lib; stream = _lib.stream_ptr(x.device)
    for kind, a in m._plan['steps']:
        if kind != 'conv': continue
        d = a['desc']
        if (d.cin, d.cout, d.ksize, d.in_h) not in want: continue
        want.remove((d.cin, d.cout, d.ksize, d.in_h))
        def run():
            lib.ryolo_conv_bn_act_fwd(ctypes.byref(d), ctypes.c_void_p(a['x']), _lib.ptr(a['w']), _lib.ptr(a['b']), ctypes.c_void_p(a['r']) if a['r'] else None, ctypes.c_void_p(a['y']), None, 0, stream)
        clk, pw, stop = [], [], False
        def sample():
            while not stop:
                clk.append(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)); pw.append(pynvml.nvmlDeviceGetPowerUsage(h) / 1000); time.sleep(0.02)
        for _ in range(50): run()
        torch.cuda.synchronize()
        t = threading.Thread(target=sample); t.start()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        N = 10000
        e0.record()
        for _ in range(N): run()
        e1.record(); torch.cuda.synchronize(); stop = True; t.join()
        half = clk[len(clk)//2:]; ph = pw[len(pw)//2:]
        print('DBG', os.environ.get('RYOLO_CONV_DEBUG','0'), (d.cin, d.cout, d.ksize, d.in_h), 'us %.1f' % (1e3 * e0.elapsed_time(e1) / N), 'clk MHz median %d min %d' % (sorted(half)[len(half)//2], min(half)), 'power W median %.0f' % sorted(ph)[len(ph)//2], 'limit', pynvml.nvmlDeviceGetEnforcedPowerLimit(h)/1000)
