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


def build_targets(model, targets, hyp):
    """targets [nt, 7] = (image, class, x, y, w, h, theta), normalised xywh.  Returns tcls, tbox, indices, av per YOLO
    layer (model/loss.py:161-258).  Like the reference, the context factor is applied to ``targets`` in place once per
    YOLO layer (harmless for the shipped context_factor = 1.0)."""
    nt = len(targets)
    tcls, tbox, indices, av, square_ious = [], [], [], [], []
    dev = targets.device
    all_ious = None
    na = 0
    t_gwha = None
    anchor_vec = None
    for i in model.yolo_layers:
        layer = model.module_list[i]
        ng, anchor_vec = layer.ng.to(dev), layer.anchor_vec.to(dev)
        targets[:, 4] += targets[:, 5] * (hyp["context_factor"] - 1)
        targets[:, 5] *= hyp["context_factor"]
        t, a = targets, []
        gwha = t[:, 4:7].clone()
        gwha[:, :-1] *= ng
        if nt:
            all_ious = torch.stack([wh_iou(x, gwha[:, :-1]) for x in anchor_vec[:, :-1]], 0)   # [na, nt]
            na = len(anchor_vec)
            a = torch.arange(na, device=dev).view((-1, 1)).repeat([1, nt]).view(-1)
            t = targets.repeat([na, 1])
            gwha = gwha.repeat([na, 1])
            square_ious.append(all_ious.view(-1))
        b, c = t[:, :2].long().t()
        gxy = t[:, 2:4] * ng
        gi, gj = gxy.long().t()
        indices.append([b, a, gj, gi])
        gxy = gxy - gxy.floor()
        t_gwha = gwha.clone()
        tbox.append(torch.cat((gxy, gwha), 1))
        av.append(anchor_vec[a] if nt else anchor_vec[:0])
        tcls.append(c)
    if nt:
        nl = len(model.yolo_layers)
        angle_offset = (t_gwha[:, -1] - anchor_vec[:, -1].view((-1, 1)).repeat([1, nt]).view(-1)).abs()
        big = angle_offset > 0.5 * math.pi
        angle_offset[big] = math.pi - angle_offset[big]
        j_a = angle_offset < model.hyp["ang_t"]
        j = [(sq > model.hyp["iou_t"]) & j_a for sq in square_ious]
        gt_j = torch.stack([juu.reshape(all_ious.shape).max(0)[0] for juu in j], 0).t()    # [nt, layers]
        # Host synchronisations are kept to a handful (one per device->host read below): the reference walks the targets
        # in a Python loop with one sync each, which on a GPU costs more than the whole loss arithmetic.
        class_ok = tcls[0].max() <= model.nc
        orphans = (~gt_j.any(1)).nonzero().view(-1).tolist()                                # sync 1
        for gt_id in orphans:       # a GT no anchor of any layer accepted: give it its best-IoU anchor (:235-242)
            gt_ious = torch.cat([sq[gt_id::nt] for sq in square_ious], 0)
            best = torch.where(gt_ious == gt_ious.max(0)[0])[0]
            layer_id = int((best // na)[0])
            best_ang = angle_offset[gt_id::nt].repeat(nl)[best].min(0)[1]
            best = best[best_ang]
            j[layer_id][(best % na) * nt + gt_id] = True
        keep = [m.nonzero().view(-1) for m in j]                                            # syncs 2..nl+1
        assert bool(class_ok), "Target classes exceed model classes"
        assert sum(len(k) for k in keep) >= nt, "something wrong at target building"
        for lid, k in enumerate(keep):
            tbox[lid] = tbox[lid][k]
            tcls[lid] = tcls[lid][k]
            av[lid] = av[lid][k]
            indices[lid] = [indices[lid][q][k] for q in range(4)]
    return tcls, tbox, indices, av


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


def compute_loss(p, targets, model, hyp, masked=True):
    """p: list of [B, na, ny, nx, nc+6] raw head tensors (training-mode output of Darknet); returns (loss[1],
    detached (lobj, lcls, lreg, loss)) exactly like model/loss.py:266-367 for arc 'default'.

    masked=True (default) evaluates the same sums over ALL (anchor, target) rows weighted by the assignment mask -- no
    host synchronisation; masked=False is the literal indexed formulation of the reference (tests compare the two)."""
    if masked and len(targets):
        return _compute_loss_masked(p, targets, model, hyp)
    return _compute_loss_indexed(p, targets, model, hyp)


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
            e = nn.functional.binary_cross_entropy_with_logits(ps[:, 5:], t, pos_weight=cls_pw, reduction="none")
            lcls = lcls + (e * mf[:, None]).sum() / (cnt * e.shape[1])
        lobj = lobj + BCEobj(pi[..., 5], tobj)
    lobj = lobj * h["obj"]
    lcls = lcls * h["cls"]
    lreg = lreg * h["reg"]
    loss = lobj + lcls + lreg
    return loss, torch.cat((lobj, lcls, lreg, loss)).detach()


def _compute_loss_indexed(p, targets, model, hyp):
    dev = p[0].device
    lcls = torch.zeros(1, device=dev)
    lobj = torch.zeros(1, device=dev)
    lreg = torch.zeros(1, device=dev)
    tcls, tbox, indices, anchor_vecs = build_targets(model, targets, hyp)
    h = model.hyp
    arc = model.arc
    if "default" not in arc or "F" in arc:
        raise NotImplementedError("only arc='default' is restated (the configuration the reference ships)")
    BCEcls = nn.BCEWithLogitsLoss(pos_weight=torch.tensor([h["cls_pw"]], device=dev))
    BCEobj = nn.BCEWithLogitsLoss(pos_weight=torch.tensor([h["obj_pw"]], device=dev))
    SM = nn.SmoothL1Loss(reduction="mean")
    for i, pi in enumerate(p):
        b, a, gj, gi = indices[i]
        tobj = torch.zeros_like(pi[..., 0])
        nb = len(b)
        if nb:
            ps = pi[b, a, gj, gi]
            tobj[b, a, gj, gi] = 1.0
            av = model.module_list[model.yolo_layers[i]].anchor_vec.to(dev)
            pxy = torch.sigmoid(ps[:, 0:2])
            pwh = torch.exp(ps[:, 2:4]).clamp(max=1e3) * av[a][:, :-1]
            pa = torch.atan(ps[:, 4]) + av[a][:, -1]
            liou = (1.0 - wh_iou(tbox[i][:, 2:4], pwh)).mean()
            lreg = lreg + SM(pxy, tbox[i][:, [0, 1]]) + 2 * SM(pa, tbox[i][:, 4]) + liou * h["giou"]
            if model.nc > 1:
                t = torch.zeros_like(ps[:, 6:])
                t[range(nb), tcls[i]] = 1.0
                lcls = lcls + BCEcls(ps[:, 5:], t)     # (reference passes ps[:, 5:] -- obj column included -- vs a [nb, nc] target
                                                      #  only when nc > 1; shapes as in the reference)
        lobj = lobj + BCEobj(pi[..., 5], tobj)         # NoSampler: the whole map (loss.py:24-28, 346-348)
    lobj = lobj * h["obj"]
    lcls = lcls * h["cls"]
    lreg = lreg * h["reg"]
    loss = lobj + lcls + lreg
    return loss, torch.cat((lobj, lcls, lreg, loss)).detach()
