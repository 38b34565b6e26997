"""Batched device-side prediction <-> ground-truth matching for mAP.

The reference (test.py:134-151) walks the predictions of an image in Python and calls ``skew_bbox_iou(pbox, tbox[m])``
once per prediction -- one kernel-equivalent plus a GPU<->CPU sync per box.  Here ONE rotated-IoU matrix launch covers
the whole image and the greedy assignment runs on the small [P, T] result."""
import torch

from .iou import rotated_iou_matrix


def match_detections(pred, tbox, tcls, iou_thres=0.5):
    """pred [P, 8] = (x, y, w, h, theta, conf, cls_conf, cls), confidence-sorted as non_max_suppression returns it;
    tbox [T, 5] pixel (x, y, w, h, theta); tcls [T].  Returns ``correct`` (list of 0/1, len P) with the reference's
    semantics: in order, a prediction is correct if its best-IoU target OF THE SAME CLASS has IoU > iou_thres and was
    not claimed before; stops early once every target is claimed (test.py:137-151)."""
    p, t = len(pred), len(tbox)
    correct = [0] * p
    if p == 0 or t == 0:
        return correct
    iou = rotated_iou_matrix(pred[:, :5].contiguous(), tbox[:, :5].contiguous().to(pred.device)).cpu()   # one launch, one copy
    pcls = pred[:, 7].cpu()
    tc = tcls.cpu().float()
    tset = set(tc.tolist())
    detected = []
    for i in range(p):
        if len(detected) == t:
            break
        if float(pcls[i]) not in tset:
            continue
        m = (tc == pcls[i]).nonzero().view(-1)
        best, bi = iou[i, m].max(0)
        if float(best) > iou_thres and int(m[bi]) not in detected:
            correct[i] = 1
            detected.append(int(m[bi]))
    return correct
