"""Batched device-side prediction <-> ground-truth matching for mAP.

The reference (test.py:134-151) walks the predictions of an image in Python and calls ``skew_bbox_iou(pbox, tbox[m])``
once per prediction -- one kernel-equivalent plus a GPU<->CPU sync per box.  Here ONE rotated-IoU matrix launch covers
the whole image and ONE small kernel does the greedy assignment (class-restricted argmax per prediction, first maximum
like torch.max, each target claimed once); the host reads the result once per image."""
import torch

from . import _lib
from .iou import rotated_iou_matrix


def match_detections_device(pred, tbox, tcls, iou_thres=0.5):
    """pred [P, 8] = (x, y, w, h, theta, conf, cls_conf, cls) CUDA, confidence-sorted as non_max_suppression returns it;
    tbox [T, 5] pixel (x, y, w, h, theta); tcls [T].  Returns ``correct`` as a CUDA uint8 tensor [P] (no host sync)."""
    p, t = len(pred), len(tbox)
    dev = pred.device
    correct = torch.zeros(p, dtype=torch.uint8, device=dev)
    if p == 0 or t == 0:
        return correct
    with torch.cuda.device(dev):
        pred = pred.float().contiguous()
        tb = tbox[:, :5].float().contiguous().to(dev)
        tc = tcls.float().contiguous().to(dev)
        iou = rotated_iou_matrix(pred[:, :5].contiguous(), tb)
        claimed = torch.empty(t, dtype=torch.uint8, device=dev)
        st = _lib.lib.ryolo_match_detections(_lib.ptr(iou), p, t, t, _lib.ptr(pred[:, 7:]), pred.shape[1], _lib.ptr(tc),
                                             float(iou_thres), _lib.ptr(claimed), _lib.ptr(correct), _lib.stream_ptr(dev))
        _lib.check(st, "ryolo_match_detections")
    return correct


def match_detections(pred, tbox, tcls, iou_thres=0.5):
    """same, as the list of 0/1 the reference builds (test.py:131-151): one host read per image"""
    return match_detections_device(pred, tbox, tcls, iou_thres).tolist()
