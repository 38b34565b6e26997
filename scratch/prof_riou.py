"""ncu target: a few launches of the 10k x 10k rotated-IoU matrix (BASELINE configs[1]).
   python scratch/prof_riou.py [reps]"""
import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
import torch
import helpers
import rotate_yolov3_b200 as pkg
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
dev = torch.device("cuda", 0)
a = helpers.gen_boxes(10000, 11, 608.0).to(dev)
b = helpers.gen_boxes(10000, 12, 608.0).to(dev)
out = pkg.rotated_iou_matrix(a, b)
torch.cuda.synchronize()
ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
ts = []
for _ in range(reps):
    flush.zero_()
    ev[0].record()
    out = pkg.rotated_iou_matrix(a, b)
    ev[1].record()
    torch.cuda.synchronize()
    ts.append(ev[0].elapsed_time(ev[1]))
print("riou 10k x 10k: ms", ["%.4f" % t for t in ts], "nonzero frac %.4f" % float((out > 0).float().mean()))
