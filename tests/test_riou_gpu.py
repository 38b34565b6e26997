"""GPU parity of ryolo_riou_* with the float64 oracle of skew_bbox_iou (tolerance from north_star: 1e-4 relative,
plus 1e-6 absolute for fp32 sliver cancellation)."""
import math
import os

import numpy as np
import pytest
import torch

from helpers import GOLDEN, gen_boxes, orc_skew_paired, orc_skew_pairwise

pytestmark = pytest.mark.gpu
RTOL, ATOL = 1e-4, 1e-6


def _close(got, ref):
    err = np.abs(got - ref)
    bad = err > RTOL * np.abs(ref) + ATOL
    assert not bad.any(), "worst offenders: %s" % sorted(zip(err[bad].tolist(), ref[bad].tolist()))[-5:]


@pytest.mark.parametrize("n,m,canvas", [(1, 1, 50), (31, 127, 100), (32, 128, 100), (33, 129, 100), (257, 515, 200),
                                        (1000, 1000, 608)])
def test_pairwise_vs_oracle(n, m, canvas):
    import rotate_yolov3_b200 as pkg
    a, b = gen_boxes(n, 11 + n, float(canvas)), gen_boxes(m, 12 + m, float(canvas))
    got = pkg.rotated_iou_matrix(a.cuda(), b.cuda()).cpu().numpy()
    _close(got, orc_skew_pairwise(a.numpy(), b.numpy(), 0))
    got_g = pkg.rotated_iou_matrix(a.cuda(), b.cuda(), GIoU=True).cpu().numpy()
    _close(got_g, orc_skew_pairwise(a.numpy(), b.numpy(), 1))


def test_paired_vs_reference_golden():
    """fixture = the reference's skew_bbox_iou python path (tests/golden/make_golden.py)"""
    import rotate_yolov3_b200 as pkg
    g = np.load(os.path.join(GOLDEN, "skew_iou_golden.npz"))
    b1, b2 = torch.from_numpy(g["b1"]).cuda(), torch.from_numpy(g["b2"]).cuda()
    out = pkg.skew_bbox_iou(b1, b2)
    assert out.is_cuda and out.dtype == torch.float32 and out.shape == (len(b2),)
    assert np.allclose(out.cpu().numpy(), g["iou"], rtol=RTOL, atol=2e-5)   # reference rounds corners to fp32
    assert np.allclose(pkg.skew_bbox_iou(b1, b2, GIoU=True).cpu().numpy(), g["giou"], rtol=RTOL, atol=2e-5)
    one = int(g["one_idx"])
    lst = [float(v) for v in g["b1"][one]]
    assert np.allclose(pkg.skew_bbox_iou(lst, b2).cpu().numpy(), g["iou_1n"], rtol=RTOL, atol=2e-5)      # list box1
    assert np.allclose(pkg.skew_bbox_iou(b1[one + 1], b2).cpu().numpy(), g["iou_1n_t"], rtol=RTOL, atol=2e-5)  # 1-D
    wide1 = torch.cat([b1, torch.rand(len(b1), 3, device="cuda")], 1)
    wide2 = torch.cat([b2, torch.rand(len(b2), 3, device="cuda")], 1)
    assert np.allclose(pkg.skew_bbox_iou(wide1, wide2).cpu().numpy(), g["iou"], rtol=RTOL, atol=2e-5)


def test_analytic_and_edge_cases():
    import rotate_yolov3_b200 as pkg

    def iou(a, b, g=False):
        return float(pkg.skew_bbox_iou(torch.tensor([a], device="cuda"), torch.tensor([b], device="cuda"), GIoU=g)[0])
    assert iou([10., 10, 8, 4, 0.3], [10., 10, 8, 4, 0.3]) == pytest.approx(1.0, abs=1e-5)
    assert iou([0., 0, 2, 2, 0], [0., 0, 2, 2, math.pi / 4]) == pytest.approx(0.7071068, rel=1e-4)
    assert iou([0., 0, 4, 2, 0], [1., 0, 4, 2, 0]) == pytest.approx(0.6, rel=1e-5)
    assert iou([0., 0, 4, 2, 0], [100., 0, 4, 2, 0.4]) == 0.0
    assert iou([0., 0, 0, 2, 0], [0., 0, 4, 2, 0]) == 0.0
    assert iou([0., 0, 4, 2, math.pi / 2], [0., 0, 2, 4, 0]) == pytest.approx(1.0, abs=1e-5)
    assert iou([0., 0, 4, 2, 0], [1., 0, 4, 2, 0], True) == pytest.approx(0.6, rel=1e-5)
    assert iou([float("nan"), 0, 4, 2, 0], [1., 0, 4, 2, 0]) == 0.0
    # far from the origin: the local-frame clip keeps fp32 cancellation at box scale
    assert iou([5e4, 7e4, 40, 10, 0.7], [5e4 + 3, 7e4 + 1, 40, 10, 0.7]) == pytest.approx(
        float(orc_skew_paired(np.float32([[5e4, 7e4, 40, 10, 0.7]]), np.float32([[5e4 + 3, 7e4 + 1, 40, 10, 0.7]]))[0]), rel=1e-3)
    empty = pkg.skew_bbox_iou(torch.zeros((0, 5), device="cuda"), torch.zeros((0, 5), device="cuda"))
    assert empty.shape == (0,)
    assert pkg.rotated_iou_matrix(torch.zeros((0, 5), device="cuda"), torch.zeros((7, 5), device="cuda")).shape == (0, 7)


def test_identical_and_touching_boxes():
    import rotate_yolov3_b200 as pkg
    a = gen_boxes(500, 3, 300.0).cuda()
    d = pkg.skew_bbox_iou(a, a.clone())
    assert float((d - 1).abs().max()) < 2e-5                      # shapely semantics: duplicates -> 1
    sh = a.clone(); sh[:, 0] += sh[:, 2] * torch.cos(sh[:, 4]); sh[:, 1] += sh[:, 2] * torch.sin(sh[:, 4])
    t = pkg.skew_bbox_iou(a, sh)                                   # sharing one short edge -> ~0
    assert float(t.max()) < 1e-4


def test_full_size_config2_properties():
    """BASELINE config 2: 10k x 10k.  Size-independent properties + an oracle-checked random sample."""
    import rotate_yolov3_b200 as pkg
    n = 10000
    a, b = gen_boxes(n, 0).cuda(), gen_boxes(n, 1).cuda()
    m = pkg.rotated_iou_matrix(a, b)
    assert m.shape == (n, n) and bool(torch.isfinite(m).all())
    assert float(m.min()) >= 0.0 and float(m.max()) <= 1.0 + 1e-6
    frac = float((m > 0).float().mean())
    assert 0.08 < frac < 0.16, frac                                 # ~12 % of pairs overlap
    mt = pkg.rotated_iou_matrix(b, a)                               # symmetry: iou(a_i,b_j) == iou(b_j,a_i)
    assert float((m - mt.t()).abs().max()) < 2e-5
    diag = pkg.rotated_iou_matrix(a, a)
    assert float((diag.diagonal() - 1).abs().max()) < 2e-5
    g = torch.Generator().manual_seed(5)
    ri, ci = torch.randint(0, n, (64,), generator=g), torch.randint(0, n, (2000,), generator=g)
    sub = m[ri.cuda()][:, ci.cuda()].cpu().numpy()
    ref = orc_skew_pairwise(a[ri.cuda()].cpu().numpy(), b[ci.cuda()].cpu().numpy())
    err = np.abs(sub - ref)
    assert np.all(err <= RTOL * np.abs(ref) + ATOL), float(err.max())
    # paired kernel agrees with the matrix kernel
    p = pkg.skew_bbox_iou(a, b)
    assert float((p - m.diagonal()).abs().max()) < 1e-6


def test_match_detections_vs_reference_loop():
    """device-batched mAP matching (metrics.match_detections) vs a direct restatement of test.py:134-151 driven by the
    float64 oracle IoU"""
    from rotate_yolov3_b200.metrics import match_detections
    g = torch.Generator().manual_seed(4)
    tb = gen_boxes(12, 50, 200.0)
    tcls = torch.randint(0, 3, (12,), generator=g).float()
    pb = torch.cat([tb[torch.randint(0, 12, (40,), generator=g)] + 3.0 * torch.randn(40, 5, generator=g) * torch.tensor([1, 1, 1, 1, 0.01]),
                    gen_boxes(20, 51, 200.0)], 0)
    conf = torch.rand(60, generator=g).sort(descending=True).values
    pcls = torch.randint(0, 4, (60,), generator=g).float()
    pred = torch.cat([pb, conf[:, None], torch.ones(60, 1), pcls[:, None]], 1)
    got = match_detections(pred.cuda(), tb.cuda(), tcls.cuda(), iou_thres=0.3)
    iou = orc_skew_pairwise(pb.numpy(), tb.numpy())
    want, detected = [0] * 60, []
    for i in range(60):
        if len(detected) == 12:
            break
        if float(pcls[i]) not in set(tcls.tolist()):
            continue
        m = (tcls == pcls[i]).nonzero().view(-1)
        row = torch.from_numpy(iou[i])[m]
        best, bi = row.max(0)
        if float(best) > 0.3 and int(m[bi]) not in detected:
            want[i] = 1
            detected.append(int(m[bi]))
    assert got == want and sum(got) > 3
