import ctypes, sys, numpy as np, torch
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import helpers, rotate_yolov3_b200 as pkg
from helpers import P
orc = helpers.oracle()
dets = helpers.adversarial_dets()
n = len(dets)
for thr in (0.0, 0.1, 0.3, 0.5, 0.7):
    keep, boxes, order, mask = pkg.nms.rnms_debug(dets.cuda(), thr)
    cb = (n+63)//64
    ref_mask = torch.zeros((n, cb), dtype=torch.int64, device='cuda')
    helpers.ref_lib('cuda').ref_cuda_mask(ctypes.c_void_p(boxes.data_ptr()), n, ctypes.c_float(thr), ctypes.c_void_p(ref_mask.data_ptr()))
    m1 = mask.cpu().numpy().view(np.uint64); m2 = ref_mask.cpu().numpy().view(np.uint64)
    bx = boxes.cpu().numpy()
    nd = 0
    for i in range(n):
        for j in range(i+1, n):
            b1 = (m1[i, j//64] >> np.uint64(j%64)) & np.uint64(1); b2 = (m2[i, j//64] >> np.uint64(j%64)) & np.uint64(1)
            if b1 != b2:
                nd += 1
                a = np.ascontiguousarray(bx[i]); b = np.ascontiguousarray(bx[j])
                v0 = orc.orc_ref_iou(P(a), P(b)); 
                v1 = orc.orc_ref_iou_fma(P(a), P(b)); npts = ctypes.c_int.in_dll(orc, 'orc_last_npts_fma').value
                if nd <= 12: print(thr, i, j, 'ours', int(b1), 'ref', int(b2), 'orc', v0, 'orc_fma', v1, 'npts', npts, a[:5], b[:5])
    print('thr', thr, 'mismatching pairs', nd)
