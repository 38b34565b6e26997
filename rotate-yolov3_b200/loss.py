"""``compute_loss`` / ``build_targets`` with the reference's semantics (model/loss.py:161-258, 266-367; wh_iou
utils/utils.py:346-361), restated device-agnostically (the reference hard-codes ``.cuda()`` at loss.py:197).  This is the
CONSUMER of the training hot path: tiny tensors, plain PyTorch on whatever device ``p`` lives on.

Covered: the shipped configuration -- arc 'default' (separate obj / cls BCE), NoSampler, reject=True target assignment
with the orphan-GT rescue.  Regression loss = SmoothL1(sigmoid xy) + 2 SmoothL1(theta) + giou * mean(1 - wh_iou): there
is no rotated IoU in the reference loss."""
import ctypes
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
    if p[0].is_cuda and p[0].dtype == torch.float32:
        return _compute_loss_fused(p, targets, model, hyp)
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


def _sparse_terms(ps, r, model, h, cls_pw):
    """regression (+ class) loss of one head over ALL (anchor, target) rows weighted by the 0/1 assignment mask; ps =
    pi[b, a, gj, gi] with the unselected rows already zeroed.  Returns (lreg_i, lcls_i), unweighted."""
    av, tbox = r["av"], r["tbox"]
    mf = r["mask"].to(ps.dtype)
    cnt = mf.sum().clamp(min=1.0)              # an empty selection contributes 0 like the reference's `if nb:`
    pxy = torch.sigmoid(ps[:, 0:2])
    pwh = torch.exp(ps[:, 2:4]).clamp(max=1e3) * av[:, :-1]
    pa = torch.atan(ps[:, 4]) + av[:, -1]
    liou = ((1.0 - wh_iou(tbox[:, 2:4], pwh)) * mf).sum() / cnt
    sm_xy = (nn.functional.smooth_l1_loss(pxy, tbox[:, 0:2], reduction="none") * mf[:, None]).sum() / (2.0 * cnt)
    sm_a = (nn.functional.smooth_l1_loss(pa, tbox[:, 4], reduction="none") * mf).sum() / cnt
    lreg = sm_xy + 2 * sm_a + liou * h["giou"]
    lcls = None
    if model.nc > 1:
        t = torch.zeros_like(ps[:, 6:])
        t[torch.arange(len(r["b"]), device=ps.device), r["tcls"]] = 1.0
        e = nn.functional.binary_cross_entropy_with_logits(ps[:, 6:], t, pos_weight=cls_pw, reduction="none")
        lcls = (e * mf[:, None]).sum() / (cnt * e.shape[1])
    return lreg, lcls


def _tobj_of(pi, r):
    mf = r["mask"].to(pi.dtype)
    tobj = torch.zeros(pi.shape[:4], dtype=pi.dtype, device=pi.device)
    tobj.index_put_((r["b"], r["a"], r["gj"], r["gi"]), mf, accumulate=True)
    return tobj.clamp_(max=1.0)


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
        tobj = _tobj_of(pi, r)
        ps = pi[r["b"], r["a"], r["gj"], r["gi"]]
        # rows the reference never touches must not poison the sums: exp(inf) * 0 would be NaN in forward AND backward
        ps = torch.where(r["mask"][:, None], ps, torch.zeros_like(ps))
        lreg_i, lcls_i = _sparse_terms(ps, r, model, h, cls_pw)
        lreg = lreg + lreg_i
        if lcls_i is not None:
            lcls = lcls + lcls_i
        lobj = lobj + BCEobj(pi[..., 5], tobj)
    lobj = lobj * h["obj"]
    lcls = lcls * h["cls"]
    lreg = lreg * h["reg"]
    loss = lobj + lcls + lreg
    return loss, torch.cat((lobj, lcls, lreg, loss)).detach()


# ---------------------------------------------------------------------------------------------------------------
# CUDA path: the dense objectness term and the whole head cotangent in two hand-written kernels (csrc/loss.cu)
# ---------------------------------------------------------------------------------------------------------------
class _FusedLoss(torch.autograd.Function):
    """loss = obj * sum_i mean BCE(pi[..., 5], tobj_i) + reg * lreg + cls * lcls with the values of _compute_loss_masked.

    forward : the matched rows are gathered (tiny [na*nt, no] tensors) and their terms -- and the gradient with respect to
              the gathered rows -- evaluated with the same framework expressions as the masked form; the objectness term
              over all 35 M cells is ONE kernel per head reading the head tensor through its strides (so the permuted
              view of the NCHW buffer the head convolution wrote is consumed in place).
    backward: ONE kernel per head writes the whole cotangent (objectness channel, zeros elsewhere) with the strides of
              the input, the row gradients are scattered on top.  Nothing reads a device value on the host."""

    @staticmethod
    def _row_args(r):
        from . import _lib
        m8 = r["mask"].to(torch.uint8)
        return (_lib.ptr(r["b"]), _lib.ptr(r["a"]), _lib.ptr(r["gj"]), _lib.ptr(r["gi"]), _lib.ptr(m8), int(r["b"].numel())), m8

    @staticmethod
    def forward(ctx, model, rows, h, *p):
        from . import _lib
        lib = _lib.lib
        dev = p[0].device
        st = _lib.stream_ptr(dev)
        cls_pw = torch.full((1,), float(h["cls_pw"]), device=dev)
        lreg = torch.zeros(1, device=dev)
        lcls = torch.zeros(1, device=dev)
        row_grads, tobjs, counts, keep = [], [], [], []
        sums = torch.zeros(len(p), dtype=torch.float64, device=dev)
        with torch.cuda.device(dev):
            gathered = []
            for i, pi in enumerate(p):
                B, na, ny, nx, no = pi.shape
                strides = (ctypes.c_longlong * 5)(*pi.stride())
                tobj = torch.zeros((B, na, ny, nx), dtype=pi.dtype, device=dev)
                if rows is not None:
                    r = rows[i]
                    for key in ("b", "a", "gj", "gi"):
                        r[key] = r[key].contiguous()
                    ra, m8 = _FusedLoss._row_args(r)
                    keep.append(m8)
                    ps = torch.empty((ra[5], no), dtype=pi.dtype, device=dev)
                    _lib.check(lib.ryolo_loss_rows_gather(_lib.ptr(pi), strides, B, na, ny, nx, no, *ra, _lib.ptr(ps), st),
                               "loss_rows_gather")
                    _lib.check(lib.ryolo_loss_rows_set_tobj(_lib.ptr(tobj), B, na, ny, nx, *ra, st), "loss_rows_set_tobj")
                    gathered.append(ps.requires_grad_())
                _lib.check(lib.ryolo_obj_bce_fwd(_lib.ptr(pi), strides, B, na, ny, nx, no, 5, _lib.ptr(tobj),
                                                 float(h["obj_pw"]), ctypes.c_void_p(sums[i:].data_ptr()), st), "obj_bce_fwd")
                tobjs.append(tobj)
                counts.append(float(B * na * ny * nx))
            if rows is not None:
                with torch.enable_grad():
                    sparse = torch.zeros(1, device=dev)
                    for ps, r in zip(gathered, rows):
                        lreg_i, lcls_i = _sparse_terms(ps, r, model, h, cls_pw)
                        lreg = lreg + lreg_i.detach()
                        sparse = sparse + lreg_i * h["reg"]
                        if lcls_i is not None:
                            lcls = lcls + lcls_i.detach()
                            sparse = sparse + lcls_i * h["cls"]
                    row_grads = [g_.contiguous() for g_ in torch.autograd.grad(sparse, gathered)]
        lobj = sums[0] * (h["obj"] / counts[0])          # python scalars only: no host-to-device copy, no sync
        for i in range(1, len(p)):
            lobj = lobj + sums[i] * (h["obj"] / counts[i])
        lobj = lobj.float().view(1)
        lreg = lreg * h["reg"]
        lcls = lcls * h["cls"]
        loss = lobj + lcls + lreg
        ctx.rows, ctx.tobjs, ctx.row_grads, ctx.counts, ctx.keep = rows, tobjs, row_grads, counts, keep
        ctx.obj_w, ctx.obj_pw = float(h["obj"]), float(h["obj_pw"])
        ctx.save_for_backward(*p)
        items = torch.cat((lobj, lcls, lreg, loss)).detach()
        ctx.mark_non_differentiable(items)
        return loss, items

    @staticmethod
    def backward(ctx, gloss, _gitems):
        from . import _lib
        lib = _lib.lib
        p = ctx.saved_tensors
        dev = p[0].device
        gloss = gloss.reshape(-1)[:1].float().contiguous()
        st = _lib.stream_ptr(dev)
        grads = []
        with torch.cuda.device(dev):
            for i, pi in enumerate(p):
                B, na, ny, nx, no = pi.shape
                # the cotangent gets the memory layout of the input: for the permuted view of an NCHW head buffer (what
                # Darknet.forward returns) it reaches the network's backward as a contiguous NCHW tensor, no copy
                g = torch.empty_strided(pi.shape, pi.stride(), dtype=pi.dtype, device=dev)   # every element is written
                scale = gloss * (ctx.obj_w / ctx.counts[i])
                strides = (ctypes.c_longlong * 5)(*pi.stride())
                _lib.check(lib.ryolo_obj_bce_bwd(_lib.ptr(pi), strides, B, na, ny, nx, no, 5, _lib.ptr(ctx.tobjs[i]),
                                                 ctx.obj_pw, _lib.ptr(scale), _lib.ptr(g), st), "obj_bce_bwd")
                if ctx.rows is not None:
                    r = ctx.rows[i]
                    ra = (_lib.ptr(r["b"]), _lib.ptr(r["a"]), _lib.ptr(r["gj"]), _lib.ptr(r["gi"]), _lib.ptr(ctx.keep[i]),
                          int(r["b"].numel()))
                    _lib.check(lib.ryolo_loss_rows_scatter_add(_lib.ptr(g), strides, B, na, ny, nx, no, *ra,
                                                               _lib.ptr(ctx.row_grads[i]), _lib.ptr(gloss), st),
                               "loss_rows_scatter_add")
                grads.append(g)
        return (None, None, None) + tuple(grads)


def _compute_loss_fused(p, targets, model, hyp):
    h = model.hyp
    if "default" not in model.arc or "F" in model.arc:
        raise NotImplementedError("only arc='default' is restated (the configuration the reference ships)")
    rows = _targets_masked(model, targets, hyp) if len(targets) else None
    return _FusedLoss.apply(model, rows, h, *p)
