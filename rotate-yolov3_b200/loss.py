"""``compute_loss`` / ``build_targets`` with the reference's semantics (model/loss.py:161-258, 266-367; wh_iou
utils/utils.py:346-361), restated device-agnostically (the reference hard-codes ``.cuda()`` at loss.py:197).  This is the
CONSUMER of the training hot path: tiny tensors, plain PyTorch on whatever device ``p`` lives on.

Covered: the shipped configuration -- arc 'default' (separate obj / cls BCE), NoSampler, reject=True target assignment
with the orphan-GT rescue.  Regression loss = SmoothL1(sigmoid xy) + 2 SmoothL1(theta) + giou * mean(1 - wh_iou): there
is no rotated IoU in the reference loss."""
import math

import torch
import torch.nn as nn


def wh_iou(box1, box2):
    """utils/utils.py:346-361"""
    if box1.shape != box2.shape:
        box2 = box2.t()
        w1, h1 = box1[0], box1[1]
        w2, h2 = box2[0], box2[1]
    else:
        w1, h1 = box1[:, 0], box1[:, 1]
        w2, h2 = box2[:, 0], box2[:, 1]
    inter = torch.min(w1, w2) * torch.min(h1, h2)
    return inter / ((w1 * h1 + 1e-16) + w2 * h2 - inter)


def _targets_masked(model, targets, hyp):
    """The target assignment of build_targets WITHOUT its data-dependent filtering: every (anchor, target) pair keeps its
    row and a boolean mask says which rows the reference would have kept.  All shapes are functions of (na, nt) only, so
    nothing here reads a device value on the host -- on a GPU the indexed formulation costs ~10 host synchronisations
    (each one drains the queue of the 75-layer forward) and leaves the device idle for longer than the loss itself takes.
    Same arithmetic as build_targets (model/loss.py:161-258), including the orphan-GT rescue (:235-242) vectorised over
    the targets."""
    nt = len(targets)
    dev = targets.device
    out = []
    square_ious = []
    anchor_vec = None
    na = 0
    t_gwha = None
    for i in model.yolo_layers:
        layer = model.module_list[i]
        ng, anchor_vec = layer.ng.to(dev), layer.anchor_vec.to(dev)
        targets[:, 4] += targets[:, 5] * (hyp["context_factor"] - 1)
        targets[:, 5] *= hyp["context_factor"]
        gwha = targets[:, 4:7].clone()
        gwha[:, :-1] *= ng
        # wh_iou of every anchor with every target in one shot (same element-wise arithmetic as the reference's
        # per-anchor loop, which on a GPU costs ~8 kernel launches per anchor)
        w1, h1 = anchor_vec[:, 0:1], anchor_vec[:, 1:2]                  # [na, 1]
        w2, h2 = gwha[:, 0].view(1, -1), gwha[:, 1].view(1, -1)          # [1, nt]
        inter = torch.min(w1, w2) * torch.min(h1, h2)
        all_ious = inter / ((w1 * h1 + 1e-16) + w2 * h2 - inter)         # [na, nt]
        na = len(anchor_vec)
        a = torch.arange(na, device=dev).view((-1, 1)).repeat([1, nt]).view(-1)
        t = targets.repeat([na, 1])
        gwha = gwha.repeat([na, 1])
        square_ious.append(all_ious.reshape(-1))
        b, c = t[:, :2].long().t()
        gxy = t[:, 2:4] * ng
        gi, gj = gxy.long().t()
        gxy = gxy - gxy.floor()
        t_gwha = gwha
        out.append(dict(b=b, a=a, gj=gj, gi=gi, tbox=torch.cat((gxy, gwha), 1), av=anchor_vec[a], tcls=c))
    nl = len(model.yolo_layers)
    angle_offset = (t_gwha[:, -1] - anchor_vec[:, -1].view((-1, 1)).repeat([1, nt]).view(-1)).abs()
    angle_offset = torch.where(angle_offset > 0.5 * math.pi, math.pi - angle_offset, angle_offset)
    j_a = angle_offset < model.hyp["ang_t"]
    j = [((sq > model.hyp["iou_t"]) & j_a).view(na, nt) for sq in square_ious]
    gt_any = torch.stack([m.any(0) for m in j], 0).any(0)                      # [nt]: some anchor of some layer accepted
    # orphan rescue: best IoU over all layers' anchors; among ties the smallest angle offset picks the ANCHOR, the first
    # tie picks the LAYER (the reference's `(best // na)[0]` next to `best[best_ang]`)
    G = torch.stack(square_ious, 0).view(nl * na, nt)
    cand = G == G.max(0)[0]
    layer_id = cand.float().argmax(0) // na                                     # argmax returns the FIRST maximum
    ang = angle_offset.view(na, nt).repeat(nl, 1)
    best = torch.where(cand, ang, torch.full_like(ang, float("inf"))).argmin(0)  # first minimum among the candidates
    anchor = best % na
    orphan = ~gt_any
    onehot = torch.arange(na, device=dev).view(-1, 1) == anchor.view(1, -1)     # [na, nt]
    for lid in range(nl):
        m = j[lid] | (onehot & (orphan & (layer_id == lid)).view(1, -1))
        out[lid]["mask"] = m.view(-1)
    return out


def compute_loss(p, targets, model, hyp):
    """p: list of [B, na, ny, nx, nc+6] raw head tensors (training-mode output of Darknet); returns (loss[1],
    detached (lobj, lcls, lreg, loss)) like model/loss.py:266-367 for arc 'default'.

    The reference selects the matched (anchor, target) rows with data-dependent indexing (build_targets, loss.py:161-258:
    ~350 device-to-host reads per step); this evaluates the same sums over ALL (anchor, target) rows weighted by the 0/1
    assignment mask -- no host synchronisation.  tests/loss_indexed.py holds the literal indexed restatement and
    tests/test_loss.py proves the two equal (value, components, gradients) and pins both to the reference's outputs."""
    if len(targets) == 0:
        return _compute_loss_no_targets(p, model)
    return _compute_loss_masked(p, targets, model, hyp)


def _compute_loss_no_targets(p, model):
    """no ground truth in the batch: only the objectness term over the whole maps remains (loss.py:346-348)"""
    dev = p[0].device
    h = model.hyp
    BCEobj = nn.BCEWithLogitsLoss(pos_weight=torch.full((1,), float(h["obj_pw"]), device=dev))
    lobj = torch.zeros(1, device=dev)
    for pi in p:
        lobj = lobj + BCEobj(pi[..., 5], torch.zeros_like(pi[..., 0]))
    lobj = lobj * h["obj"]
    zero = torch.zeros(1, device=dev)
    return lobj + zero, torch.cat((lobj, zero, zero, lobj)).detach()


def _compute_loss_masked(p, targets, model, hyp):
    dev = p[0].device
    h = model.hyp
    if "default" not in model.arc or "F" in model.arc:
        raise NotImplementedError("only arc='default' is restated (the configuration the reference ships)")
    lcls = torch.zeros(1, device=dev)
    lobj = torch.zeros(1, device=dev)
    lreg = torch.zeros(1, device=dev)
    rows = _targets_masked(model, targets, hyp)
    # torch.full, not torch.tensor([...], device=...): the latter is a (synchronising) pageable host-to-device copy
    BCEobj = nn.BCEWithLogitsLoss(pos_weight=torch.full((1,), float(h["obj_pw"]), device=dev))
    cls_pw = torch.full((1,), float(h["cls_pw"]), device=dev)
    for pi, r in zip(p, rows):
        b, a, gj, gi, av, tbox = r["b"], r["a"], r["gj"], r["gi"], r["av"], r["tbox"]
        mf = r["mask"].to(pi.dtype)
        cnt = mf.sum().clamp(min=1.0)              # an empty selection contributes 0 like the reference's `if nb:`
        tobj = torch.zeros_like(pi[..., 0])
        tobj.index_put_((b, a, gj, gi), mf, accumulate=True)
        tobj.clamp_(max=1.0)
        ps = pi[b, a, gj, gi]
        # rows the reference never touches must not poison the sums: exp(inf) * 0 would be NaN in forward AND backward
        ps = torch.where(r["mask"][:, None], ps, torch.zeros_like(ps))
        pxy = torch.sigmoid(ps[:, 0:2])
        pwh = torch.exp(ps[:, 2:4]).clamp(max=1e3) * av[:, :-1]
        pa = torch.atan(ps[:, 4]) + av[:, -1]
        liou = ((1.0 - wh_iou(tbox[:, 2:4], pwh)) * mf).sum() / cnt
        sm_xy = (nn.functional.smooth_l1_loss(pxy, tbox[:, 0:2], reduction="none") * mf[:, None]).sum() / (2.0 * cnt)
        sm_a = (nn.functional.smooth_l1_loss(pa, tbox[:, 4], reduction="none") * mf).sum() / cnt
        lreg = lreg + sm_xy + 2 * sm_a + liou * h["giou"]
        if model.nc > 1:
            t = torch.zeros_like(ps[:, 6:])
            t[torch.arange(len(b), device=dev), r["tcls"]] = 1.0
            e = nn.functional.binary_cross_entropy_with_logits(ps[:, 6:], t, pos_weight=cls_pw, reduction="none")
            lcls = lcls + (e * mf[:, None]).sum() / (cnt * e.shape[1])
        lobj = lobj + BCEobj(pi[..., 5], tobj)
    lobj = lobj * h["obj"]
    lcls = lcls * h["cls"]
    lreg = lreg * h["reg"]
    loss = lobj + lcls + lreg
    return loss, torch.cat((lobj, lcls, lreg, loss)).detach()
