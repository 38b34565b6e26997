#!/usr/bin/env python
"""Generates the committed golden fixtures under tests/golden/ by RUNNING THE REFERENCE in this container
(/root/reference, read-only).  Not run by the test-suite; the GPU box has no /root/reference and only sees the
fixtures.  Re-run: `python tests/golden/make_golden.py` (after `__graft_entry__.build()`, which builds
oracle/_ref from the reference sources).

What comes from where:
  rnms_ref_golden.npz      reference device code (utils/nms/src/rotate_polygon_nms_kernel.cu:19-260) compiled as
                           host C++ by oracle/build_ref.sh: paired IoUs + r_nms keep lists (thr .1/.3/.5) + the
                           4-box fixture of utils/nms/nms_wrapper_test.py:35-38.
  rotated_coors_golden.npz reference get_rotated_coors (utils/utils.py:702-725, via cv2) on random boxes.
  skew_iou_golden.npz      reference skew_bbox_iou / skewiou PYTHON CODE (utils/utils.py:290-320, 663-699)
                           executed with a stand-in for shapely (absent from this image): a float64 convex
                           polygon class written here.  Pins broadcast rules, corner convention, zero guards and
                           the 'giou' envelope definition; the GEOS arithmetic itself stays unpinned.
  nms_driver_golden.npz    reference non_max_suppression (utils/nms/nms.py:4-69) with its r_nms extension
                           replaced by the host build of the reference kernel code.
  decode_golden.npz        reference YOLOLayer.forward eval branch + create_grids (model/models.py:183-227,
                           model/model_utils.py:16-35) on CPU.
"""
import ctypes
import math
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"


# ----------------------------------------------------------------------------------------------------------
# float64 convex-polygon stand-in for the subset of shapely the reference touches
# ----------------------------------------------------------------------------------------------------------
def _hull(pts):
    pts = sorted(set((float(x), float(y)) for x, y in pts))
    if len(pts) <= 2:
        return pts

    def cross(o, a, b):
        return (a[0] - o[0]) * (b[1] - o[1]) - (a[1] - o[1]) * (b[0] - o[0])
    lo, up = [], []
    for p in pts:
        while len(lo) >= 2 and cross(lo[-2], lo[-1], p) <= 0:
            lo.pop()
        lo.append(p)
    for p in reversed(pts):
        while len(up) >= 2 and cross(up[-2], up[-1], p) <= 0:
            up.pop()
        up.append(p)
    return lo[:-1] + up[:-1]  # CCW


def _area(poly):
    n = len(poly)
    if n < 3:
        return 0.0
    return 0.5 * abs(sum(poly[i][0] * poly[(i + 1) % n][1] - poly[(i + 1) % n][0] * poly[i][1] for i in range(n)))


def _clip(subj, clip):
    out = list(subj)
    n = len(clip)
    for e in range(n):
        a, b = clip[e], clip[(e + 1) % n]
        inp, out = out, []
        if not inp:
            break
        for i in range(len(inp)):
            p, q = inp[i], inp[(i + 1) % len(inp)]
            dp = (b[0] - a[0]) * (p[1] - a[1]) - (b[1] - a[1]) * (p[0] - a[0])
            dq = (b[0] - a[0]) * (q[1] - a[1]) - (b[1] - a[1]) * (q[0] - a[0])
            if dp >= 0:
                out.append(p)
            if (dp >= 0) != (dq >= 0):
                t = dp / (dp - dq)
                out.append((p[0] + t * (q[0] - p[0]), p[1] + t * (q[1] - p[1])))
    return out


class Polygon:
    def __init__(self, pts):
        self.pts = [tuple(map(float, p)) for p in (pts.pts if isinstance(pts, Polygon) else np.asarray(pts).reshape(-1, 2))]

    @property
    def convex_hull(self):
        return Polygon(_hull(self.pts))

    @property
    def is_valid(self):
        return all(math.isfinite(v) for p in self.pts for v in p)

    @property
    def area(self):
        return _area(self.pts)

    def intersection(self, other):
        if len(self.pts) < 3 or len(other.pts) < 3:
            return Polygon(np.zeros((0, 2)))
        return Polygon(np.asarray(_clip(_hull(self.pts), _hull(other.pts))).reshape(-1, 2))


class MultiPoint(Polygon):
    @property
    def envelope(self):
        xs = [p[0] for p in self.pts]
        ys = [p[1] for p in self.pts]
        e = Polygon([(min(xs), min(ys)), (max(xs), min(ys)), (max(xs), max(ys)), (min(xs), max(ys))])
        e.wkt = "stub"
        return e

    @property
    def convex_hull(self):
        h = Polygon(_hull(self.pts))
        h.wkt = "stub"
        return h


def import_reference():
    """sys.modules stubs for what this image lacks, then import the reference."""
    mpl = types.ModuleType("matplotlib")
    mpl.rc = lambda *a, **k: None
    plt = types.ModuleType("matplotlib.pyplot")
    mpl.pyplot = plt
    sh = types.ModuleType("shapely")
    geo = types.ModuleType("shapely.geometry")
    geo.Polygon = Polygon
    geo.MultiPoint = MultiPoint
    sh.geometry = geo
    rn = types.ModuleType("utils.nms.r_nms")
    rn.r_nms = None  # filled by the caller
    sys.modules.update({"matplotlib": mpl, "matplotlib.pyplot": plt, "shapely": sh, "shapely.geometry": geo,
                        "utils.nms.r_nms": rn})
    sys.path.insert(0, REF)
    os.chdir(REF)
    torch.cuda.FloatTensor = torch.FloatTensor  # utils/utils.py:291 hard-codes the CUDA tensor type
    import cv2
    _grm = cv2.getRotationMatrix2D  # cv2 >= 4.5 no longer accepts 0-d tensors where the reference passes them (:708)
    cv2.getRotationMatrix2D = lambda angle, center, scale: _grm(angle=float(angle), center=(float(center[0]), float(center[1])), scale=float(scale))
    import utils.utils as ru  # noqa
    import utils.nms.nms as rnms  # noqa
    import model.models as rmodels  # noqa
    return ru, rnms, rmodels, rn


def gen_boxes(n, seed, canvas=608.0):
    """SURVEY.md 8(d) config-2 generator: HRSC-anchor-like boxes on a canvas^2 image."""
    g = torch.Generator().manual_seed(seed)
    cx = torch.rand(n, generator=g) * canvas
    cy = torch.rand(n, generator=g) * canvas
    area = 792 + torch.rand(n, generator=g) * (15803 - 792)
    ratio = 4 + torch.rand(n, generator=g) * 5
    w = (area * ratio).sqrt()
    h = (area / ratio).sqrt()
    th = (torch.rand(n, generator=g) - 0.5) * math.pi
    return torch.stack([cx, cy, w, h, th], 1).float()


def tie_free_scores(n, seed):
    g = torch.Generator().manual_seed(seed)
    return torch.linspace(0.01, 1, n)[torch.randperm(n, generator=g)].float()


def main():
    ref_host = ctypes.CDLL(os.path.join(REPO, "oracle", "_ref", "libref_rnms_host.so"))
    fp = ctypes.POINTER(ctypes.c_float)
    ip = ctypes.POINTER(ctypes.c_int64)

    def P(a):
        return a.ctypes.data_as(fp)

    # ---- 1. reference kernel code (host build) ----
    n = 4000
    a = np.concatenate([gen_boxes(n, 10, 200.0).numpy(), np.zeros((n, 1), np.float32)], 1)
    b = np.concatenate([gen_boxes(n, 11, 200.0).numpy(), np.zeros((n, 1), np.float32)], 1)
    a[:200] = b[:200]                                   # exact duplicates
    b[200:400, :5] = a[200:400, :5]; b[200:400, 0] += a[200:400, 2]   # same box shifted by w along x
    a[400:420, 2] = 0.0                                 # zero width
    b[420:440, 3] = 0.0
    a[440:460, 4] = 0.0; b[440:460, 4] = 0.0            # axis-aligned pairs
    b[460:480, :2] = a[460:480, :2]                     # concentric
    iou = np.empty(n, np.float32)
    ref_host.ref_host_iou_paired(P(a), P(b), n, 6, P(iou), 1)
    fx = np.array([[150, 150, 100, 100, 0, 0.99], [160, 160, 100, 100, 0, 0.88], [150, 150, 100, 100, -0.7854, 0.66],
                   [300, 300, 100, 100, 0., 0.77]], dtype=np.float32)  # utils/nms/nms_wrapper_test.py:35-38
    fx_iou = np.empty(3, np.float32)
    ref_host.ref_host_iou.restype = ctypes.c_float
    for k in range(3):
        fx_iou[k] = ref_host.ref_host_iou(P(fx[0]), P(fx[k + 1]))
    nd = 1500
    dets = np.concatenate([gen_boxes(nd, 12, 300.0).numpy(), tie_free_scores(nd, 13).numpy()[:, None]], 1).astype(np.float32)
    keeps = {}
    for thr in (0.1, 0.3, 0.5):
        k = np.empty(nd, np.int64)
        cnt = ref_host.ref_host_rnms(P(dets), nd, ctypes.c_float(thr), k.ctypes.data_as(ip), 8, None, None)
        keeps["keep_%02d" % int(thr * 10)] = k[:cnt].copy()
    kf = np.empty(4, np.int64)
    cnt = ref_host.ref_host_rnms(P(fx), 4, ctypes.c_float(0.1), kf.ctypes.data_as(ip), 1, None, None)
    np.savez_compressed(os.path.join(HERE, "rnms_ref_golden.npz"), a=a, b=b, iou=iou, fixture=fx, fixture_iou=fx_iou,
                        fixture_keep=kf[:cnt].copy(), dets=dets, **keeps)
    print("rnms_ref_golden: nonzero frac %.3f, fixture keep %s, K=%s" % ((iou > 0).mean(), kf[:cnt],
                                                                         {k: len(v) for k, v in keeps.items()}))

    # ---- reference python modules ----
    ru, rnms, rmodels, rn_stub = import_reference()

    # ---- 2. corner convention ----
    bx = gen_boxes(64, 20).numpy().astype(np.float64)
    coors = np.stack([ru.get_rotated_coors(bb) for bb in bx])
    np.savez_compressed(os.path.join(HERE, "rotated_coors_golden.npz"), boxes=bx, coors=coors)

    # ---- 3. skew_bbox_iou python path with the shapely stand-in ----
    m = 300
    b1 = gen_boxes(m, 21, 150.0)
    b2 = gen_boxes(m, 22, 150.0)
    b2[:10] = b1[:10]                 # identical -> 1
    b2[10:20, 2] = 0.0                # zero area -> 0
    b1[20:30, :2] += 5000.0           # disjoint -> 0
    out_nn = ru.skew_bbox_iou(b1, b2).numpy()
    out_nn_g = ru.skew_bbox_iou(b1, b2, GIoU=True).numpy()
    one = [float(v) for v in b1[40]]
    out_1n = ru.skew_bbox_iou(one, b2).numpy()              # list box1 (utils/utils.py:292)
    out_1n_t = ru.skew_bbox_iou(b1[41], b2).numpy()         # 1-D tensor box1 (:294-295)
    wide = torch.cat([b2, torch.rand(m, 3)], 1)             # extra columns are ignored (:299-300)
    out_wide = ru.skew_bbox_iou(b1, wide[:, :]).numpy() if False else ru.skew_bbox_iou(torch.cat([b1, torch.rand(m, 3)], 1), wide).numpy()
    np.savez_compressed(os.path.join(HERE, "skew_iou_golden.npz"), b1=b1.numpy(), b2=b2.numpy(), iou=out_nn,
                        giou=out_nn_g, one_idx=40, iou_1n=out_1n, iou_1n_t=out_1n_t, iou_wide=out_wide)
    print("skew golden: mean iou %.4f nonzero %.3f ident %s" % (out_nn.mean(), (out_nn > 0).mean(), out_nn[:10]))

    # ---- 4. non_max_suppression driver with the reference kernel code behind r_nms ----
    def r_nms_host(dets_t, thr):
        d = np.ascontiguousarray(dets_t.detach().cpu().numpy().astype(np.float32))
        if d.size == 0:
            return torch.empty((0,), dtype=torch.long)
        k = np.empty(len(d), np.int64)
        c = ref_host.ref_host_rnms(P(d), len(d), ctypes.c_float(float(thr)), k.ctypes.data_as(ip), 1, None, None)
        return torch.from_numpy(k[:c].copy())
    rnms.r_nms = r_nms_host
    g = torch.Generator().manual_seed(30)
    bs, pn, nc = 3, 600, 3
    pred = torch.zeros(bs, pn, 6 + nc)
    for i in range(bs):
        pred[i, :, :5] = gen_boxes(pn, 31 + i, 250.0)
    pred[..., 5] = torch.rand(bs, pn, generator=g)
    pred[..., 6:] = torch.rand(bs, pn, nc, generator=g)
    pred[0, :20, 2] = 1.5                 # too small (min_wh, nms.py:12,40)
    pred[0, 20:25, 0] = float("nan")      # non-finite rows dropped
    pred[0, 25:28, 7] = float("inf")
    pred[2, :, 5] = 0.01                  # image with no survivor -> None
    pred_in = pred.clone()
    out = rnms.non_max_suppression(pred, conf_thres=0.3, nms_thres=0.4)
    sav = {"pred": pred_in.numpy(), "pred_after": pred.numpy(), "conf_thres": 0.3, "nms_thres": 0.4}
    for i, o in enumerate(out):
        sav["out_%d" % i] = np.zeros((0, 8), np.float32) if o is None else o.numpy()
        sav["none_%d" % i] = o is None
    # single-class variant (nc = 1, the shipped configs)
    pred1 = torch.zeros(2, 500, 7)
    for i in range(2):
        pred1[i, :, :5] = gen_boxes(500, 41 + i, 200.0)
    pred1[..., 5] = torch.rand(2, 500, generator=g)
    pred1[..., 6] = 1.0
    sav["pred1"] = pred1.clone().numpy()
    out1 = rnms.non_max_suppression(pred1, conf_thres=0.5, nms_thres=0.5)
    for i, o in enumerate(out1):
        sav["out1_%d" % i] = np.zeros((0, 8), np.float32) if o is None else o.numpy()
    np.savez_compressed(os.path.join(HERE, "nms_driver_golden.npz"), **sav)
    print("nms driver golden:", [None if o is None else tuple(o.shape) for o in out], [tuple(o.shape) for o in out1])

    # ---- 5. YOLO decode ----
    anchors = np.array([[30.0, 10.0, -0.5], [60.0, 15.0, 0.0], [90.0, 20.0, 0.7], [45.0, 45.0, 1.2]])
    dec = {}
    for tag, nc_, ctx in (("a", 1, 1.0), ("b", 3, 1.25)):
        hyp = {"context_factor": ctx}
        layer = rmodels.YOLOLayer(anchors=anchors.copy(), nc=nc_, yolo_index=0, arc="default", hyp=hyp)
        layer.eval()
        g = torch.Generator().manual_seed(50)
        ny, nx = 5, 7
        p = torch.randn(2, len(anchors) * (nc_ + 6), ny, nx, generator=g)
        io, pp = layer(p, img_size=torch.Size([ny * 16, nx * 16]))
        dec.update({"p_" + tag: p.numpy(), "io_" + tag: io.numpy(), "pp_" + tag: pp.numpy(), "nc_" + tag: nc_,
                    "ctx_" + tag: ctx, "stride_" + tag: float(layer.stride)})
    dec["anchors"] = anchors
    np.savez_compressed(os.path.join(HERE, "decode_golden.npz"), **dec)
    print("decode golden:", {k: v.shape for k, v in dec.items() if hasattr(v, "shape") and v.ndim > 1})

    # ---- 6. whole-network eval forward of the reference Darknet (Darknet-53 graph of cfg/yolov3.cfg, 6 anchors) ----
    import tempfile
    sys.path.insert(0, os.path.join(REPO, "tests"))
    sys.path.insert(0, REPO)
    import helpers
    spec = __import__("importlib.util").util.spec_from_file_location("cfgs", os.path.join(REPO, "rotate-yolov3_b200", "cfgs.py"))
    cfgs = __import__("importlib.util").util.module_from_spec(spec)
    spec.loader.exec_module(cfgs)
    text = cfgs.yolov3_cfg(width=96, height=64, classes=1, anchors=helpers.SMALL_ANCHORS, n_anchors=6)
    with tempfile.NamedTemporaryFile("w", suffix=".cfg", delete=False) as f:
        f.write(text)
        cfg_path = f.name
    model = rmodels.Darknet(cfg_path, {"context_factor": 1.0}, arc="default")
    helpers.init_darknet_weights(model, seed=123)
    model.eval()
    g = torch.Generator().manual_seed(7)
    x = torch.rand(2, 3, 64, 96, generator=g)
    with torch.no_grad():
        io, ps = model(x)
    np.savez_compressed(os.path.join(HERE, "darknet_golden.npz"), x=x.numpy(), io=io.numpy(),
                        p0=ps[0].numpy(), p1=ps[1].numpy(), p2=ps[2].numpy(), n_params=sum(p.numel() for p in model.parameters()))
    print("darknet golden:", tuple(io.shape), [tuple(p.shape) for p in ps], float(io.abs().max()),
          [float(p.abs().max()) for p in ps])

    # ---- 7. training-mode forward + backward of the reference Darknet (batch-statistics BN, autograd) ----
    text = cfgs.yolov3_cfg(width=160, height=128, classes=1, anchors=helpers.SMALL_ANCHORS, n_anchors=6)
    with tempfile.NamedTemporaryFile("w", suffix=".cfg", delete=False) as f:
        f.write(text)
        cfg_path = f.name
    model = rmodels.Darknet(cfg_path, {"context_factor": 1.0}, arc="default")
    helpers.init_darknet_weights(model, seed=321)
    model.train()
    g = torch.Generator().manual_seed(11)
    x = torch.rand(4, 3, 128, 160, generator=g)
    ps = model(x)
    gs = [torch.randn(p.shape, generator=g) for p in ps]
    loss = sum((p * gg).sum() for p, gg in zip(ps, gs)) / 100.0
    loss.backward()
    sav = {"x": x.numpy(), "loss": float(loss)}
    for k_, (p, gg) in enumerate(zip(ps, gs)):
        sav["p%d" % k_] = p.detach().numpy()
        sav["g%d" % k_] = gg.numpy()
    names, norms, samples, sample_idx = [], [], [], []
    gi = torch.Generator().manual_seed(12)
    for name, prm in model.named_parameters():
        gflat = prm.grad.reshape(-1)
        idx = torch.randint(0, gflat.numel(), (min(256, gflat.numel()),), generator=gi)
        names.append(name)
        norms.append(float(gflat.norm()))
        sample_idx.append(idx.numpy())
        samples.append(gflat[idx].numpy())
    sav["names"] = np.array(names)
    sav["norms"] = np.array(norms)
    sav["sample_idx"] = np.array(sample_idx, dtype=object)
    sav["samples"] = np.array(samples, dtype=object)
    sav["rm0"] = model.module_list[0].BatchNorm2d.running_mean.numpy()
    sav["rv0"] = model.module_list[0].BatchNorm2d.running_var.numpy()
    sav["rm75"] = model.module_list[75].BatchNorm2d.running_mean.numpy()
    np.savez_compressed(os.path.join(HERE, "darknet_train_golden.npz"), **sav)

    # ---- 8. the same on the 19-block "mini" graph (tight wiring check: few layers, little bf16 noise) ----
    with tempfile.NamedTemporaryFile("w", suffix=".cfg", delete=False) as f:
        f.write(helpers.mini_cfg(64, 48))
        cfg_path = f.name
    mm = rmodels.Darknet(cfg_path, {"context_factor": 1.0}, arc="default")
    helpers.init_darknet_weights(mm, seed=77)
    g = torch.Generator().manual_seed(21)
    xm = torch.rand(4, 3, 48, 64, generator=g)
    mm.eval()
    with torch.no_grad():
        io_e, ps_e = mm(xm)
    mm.train()
    psm = mm(xm)
    gsm = [torch.randn(p.shape, generator=g) for p in psm]
    lm = sum((p * gg).sum() for p, gg in zip(psm, gsm)) / 10.0
    lm.backward()
    savm = {"x": xm.numpy(), "io_eval": io_e.numpy(), "loss": float(lm)}
    for k_, (p, gg, pe) in enumerate(zip(psm, gsm, ps_e)):
        savm["p%d" % k_] = p.detach().numpy()
        savm["g%d" % k_] = gg.numpy()
        savm["pe%d" % k_] = pe.numpy()
    for name, prm in mm.named_parameters():
        savm["grad:" + name] = prm.grad.numpy()
    np.savez_compressed(os.path.join(HERE, "mini_train_golden.npz"), **savm)

    # ---- 9. reference compute_loss / build_targets (model/loss.py) on the section-7 model's training output ----
    import model.loss as rloss
    torch.Tensor.cuda = lambda self, *a, **k: self          # loss.py:197 hard-codes .cuda()
    hyp_l = {"giou": 0.1, "cls": 27.76, "cls_pw": 1.0, "obj": 20.35, "obj_pw": 1.0, "iou_t": 0.5, "ang_t": 3.1415926 / 12,
             "reg": 1.0, "context_factor": 1.0}
    model.hyp, model.nc, model.arc = hyp_l, 1, "default"
    g = torch.Generator().manual_seed(31)
    nt = 9
    tg = torch.zeros(nt, 7)
    tg[:, 0] = torch.randint(0, 4, (nt,), generator=g).float()
    tg[:, 2:4] = 0.1 + 0.8 * torch.rand(nt, 2, generator=g)
    # sizes near the area-900 anchors (67 x 13.4 px on the 160 x 128 image) so that every GT is matched: the orphan-GT
    # rescue (loss.py:235-242) indexes with a true-divided tensor, which the installed torch rejects
    tg[:, 4] = (67.08 / 160.0) * (0.9 + 0.2 * torch.rand(nt, generator=g))
    tg[:, 5] = (13.42 / 128.0) * (0.9 + 0.2 * torch.rand(nt, generator=g))
    tg[:, 6] = torch.tensor([-0.8, -0.7, 0.75, 0.9, 0.6, -0.95, 0.8, -0.6, 0.7])[:nt]
    pl = [p.detach().clone().requires_grad_(True) for p in ps]
    loss_l, items = rloss.compute_loss(pl, tg.clone(), model, hyp_l)
    loss_l.backward()
    savl = {"targets": tg.numpy(), "loss": loss_l.detach().numpy(), "items": items.numpy(), "hyp_keys": np.array(list(hyp_l)),
            "hyp_vals": np.array([hyp_l[k] for k in hyp_l])}
    for k_, p in enumerate(pl):
        savl["p%d" % k_] = p.detach().numpy()
        savl["dp%d" % k_] = p.grad.numpy()
    np.savez_compressed(os.path.join(HERE, "loss_golden.npz"), **savl)
    print("loss golden:", items.numpy())

    # ---- 10. Darknet .weights written by the reference's save_weights (model/model_utils.py:96-118) ----
    import model.model_utils as rmu
    micro = ("[net]\nwidth=32\nheight=32\nchannels=3\n\n[convolutional]\nbatch_normalize=1\nfilters=8\nsize=3\nstride=1\npad=1\n"
             "activation=leaky\n\n[convolutional]\nbatch_normalize=1\nfilters=16\nsize=3\nstride=2\npad=1\nactivation=leaky\n\n"
             "[convolutional]\nfilters=14\nsize=1\nstride=1\npad=1\nactivation=linear\n\n[yolo]\nmask = 0-1\nanchors = ara 900 / 5.0 / -45, 45\n"
             "classes=1\nnum=2\n\n")
    with tempfile.NamedTemporaryFile("w", suffix=".cfg", delete=False) as f:
        f.write(micro)
        cfg_path = f.name
    mw = rmodels.Darknet(cfg_path, {"context_factor": 1.0}, arc="default")
    helpers.init_darknet_weights(mw, seed=9)
    mw.seen = np.array([1234], dtype=np.int64)
    wpath = os.path.join(HERE, "micro_reference.weights")
    rmu.save_weights(mw, wpath)
    with open(os.path.join(HERE, "micro.cfg"), "w") as f:
        f.write(micro)
    print("weights golden:", os.path.getsize(wpath), "bytes")
    print("mini golden: loss %.4f, %d grads" % (float(lm), len([k for k in savm if k.startswith("grad:")])))
    print("train golden: loss %.4f, %d params, grad norm range %.3e .. %.3e" % (float(loss), len(names), min(norms), max(norms)))


if __name__ == "__main__":
    main()
