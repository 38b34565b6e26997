"""Pins the CPU oracle (oracle/rbox_oracle.c) before anything trusts it: against the committed golden vectors
produced by the reference itself (tests/golden/make_golden.py), against the reference's own 4-box fixture
(utils/nms/nms_wrapper_test.py:35-38), against oracle/_ref when it is present, and against analytic cases."""
import ctypes
import math
import os

import numpy as np
import pytest

from helpers import GOLDEN, I64P, P, gen_boxes, gen_dets, oracle, orc_rnms, orc_skew_paired, orc_skew_pairwise, ref_lib


def test_ref_iou_bit_exact_vs_golden():
    g = np.load(os.path.join(GOLDEN, "rnms_ref_golden.npz"))
    a, b = np.ascontiguousarray(g["a"]), np.ascontiguousarray(g["b"])
    out = np.empty(len(a), np.float32)
    oracle().orc_ref_iou_paired(P(a), P(b), len(a), 6, P(out))
    assert np.array_equal(out.view(np.uint32), g["iou"].view(np.uint32))  # bit-for-bit incl. NaN payloads
    assert (g["iou"] > 0).mean() > 0.3


def test_wrapper_fixture_analytic():
    g = np.load(os.path.join(GOLDEN, "rnms_ref_golden.npz"))
    fx = np.ascontiguousarray(g["fixture"])
    ious = [oracle().orc_ref_iou(P(fx[0]), P(fx[k])) for k in (1, 2, 3)]
    assert abs(ious[0] - 8100.0 / 11900.0) < 1e-6
    assert abs(ious[1] - 2 * (math.sqrt(2) - 1) / (2 - 2 * (math.sqrt(2) - 1))) < 2e-6
    assert ious[2] == 0.0
    assert np.array_equal(np.float32(ious), g["fixture_iou"])
    assert list(orc_rnms(fx, 0.1)) == [0, 3] == list(g["fixture_keep"])


@pytest.mark.parametrize("thr", [0.1, 0.3, 0.5])
def test_rnms_keep_lists_vs_golden(thr):
    g = np.load(os.path.join(GOLDEN, "rnms_ref_golden.npz"))
    want = g["keep_%02d" % int(thr * 10)]
    got = orc_rnms(g["dets"], thr)
    assert np.array_equal(got, want)
    # the nvcc-contracted arithmetic variant takes the same decisions on these continuous inputs
    assert np.array_equal(orc_rnms(g["dets"], thr, variant=1), want)


def test_live_reference_build_when_present():
    ref = ref_lib("host")
    if ref is None:
        pytest.skip("oracle/_ref not built (no /root/reference on this box)")
    a = np.concatenate([gen_boxes(20000, 3, 150.0).numpy(), np.zeros((20000, 1), np.float32)], 1)
    b = np.concatenate([gen_boxes(20000, 4, 150.0).numpy(), np.zeros((20000, 1), np.float32)], 1)
    b[:500] = a[:500]
    o1, o2 = np.empty(20000, np.float32), np.empty(20000, np.float32)
    oracle().orc_ref_iou_paired(P(a), P(b), 20000, 6, P(o1))
    ref.ref_host_iou_paired(P(a), P(b), 20000, 6, P(o2), 4)
    assert np.array_equal(o1.view(np.uint32), o2.view(np.uint32))
    d = gen_dets(2500, 5, 300.0).numpy()
    k = np.empty(2500, np.int64)
    c = ref.ref_host_rnms(P(d), 2500, ctypes.c_float(0.4), k.ctypes.data_as(I64P), 4, None, None)
    assert np.array_equal(orc_rnms(d, 0.4), k[:c])


def test_rotated_coors_convention_vs_reference():
    g = np.load(os.path.join(GOLDEN, "rotated_coors_golden.npz"))
    out = np.empty(8, np.float64)
    for box, want in zip(g["boxes"], g["coors"]):
        bx = np.ascontiguousarray(box, dtype=np.float64)
        oracle().orc_rotated_coors(bx.ctypes.data_as(ctypes.POINTER(ctypes.c_double)),
                                   out.ctypes.data_as(ctypes.POINTER(ctypes.c_double)))
        assert np.allclose(out, want, rtol=0, atol=1e-9 * max(1.0, np.abs(want).max()))


def test_skew_iou_vs_reference_python_path():
    g = np.load(os.path.join(GOLDEN, "skew_iou_golden.npz"))
    b1, b2 = g["b1"], g["b2"]
    # the reference rounds corners to fp32 before the polygon code (tensor * float64 -> fp32): 2e-5 abs
    assert np.allclose(orc_skew_paired(b1, b2, 0), g["iou"], rtol=1e-4, atol=2e-5)
    assert np.allclose(orc_skew_paired(b1, b2, 1), g["giou"], rtol=1e-4, atol=2e-5)
    one = int(g["one_idx"])
    rep = np.repeat(b1[one:one + 1], len(b2), 0)
    assert np.allclose(orc_skew_paired(rep, b2, 0), g["iou_1n"], rtol=1e-4, atol=2e-5)
    assert np.all(g["iou"][:10] > 0.9999) and np.all(g["iou"][10:30] == 0)


def test_skew_iou_analytic():
    def iou(a, b, mode=0):
        return float(orc_skew_pairwise(np.float32([a]), np.float32([b]), mode)[0, 0])
    assert iou([10, 10, 8, 4, 0.3], [10, 10, 8, 4, 0.3]) == pytest.approx(1.0, abs=1e-6)
    assert iou([0, 0, 2, 2, 0], [0, 0, 2, 2, math.pi / 4]) == pytest.approx(
        2 * (math.sqrt(2) - 1) / (2 - 2 * (math.sqrt(2) - 1)), rel=1e-6)   # 45-degree square in square
    assert iou([0, 0, 4, 2, 0], [1, 0, 4, 2, 0]) == pytest.approx(6.0 / 10.0, rel=1e-6)  # axis-aligned offset
    assert iou([0, 0, 4, 2, 0], [100, 0, 4, 2, 0.4]) == 0.0
    assert iou([0, 0, 0, 2, 0], [0, 0, 4, 2, 0]) == 0.0                                   # zero area
    assert iou([0, 0, 4, 2, math.pi / 2], [0, 0, 2, 4, 0]) == pytest.approx(1.0, abs=1e-6)  # same rectangle
    # 'giou' of the reference = inter / envelope area
    assert iou([0, 0, 4, 2, 0], [1, 0, 4, 2, 0], 1) == pytest.approx(6.0 / 10.0, rel=1e-6)
    assert iou([0, 0, 2, 2, 0], [3, 3, 2, 2, 0], 1) == 0.0


def test_skew_iou_vs_opencv():
    cv2 = pytest.importorskip("cv2")
    a = gen_boxes(400, 8, 120.0).numpy()
    b = gen_boxes(400, 9, 120.0).numpy()
    got = orc_skew_paired(a, b, 0)
    for i in range(400):
        ra = ((float(a[i, 0]), float(a[i, 1])), (float(a[i, 2]), float(a[i, 3])), math.degrees(float(a[i, 4])))
        rb = ((float(b[i, 0]), float(b[i, 1])), (float(b[i, 2]), float(b[i, 3])), math.degrees(float(b[i, 4])))
        rc, pts = cv2.rotatedRectangleIntersection(ra, rb)
        inter = cv2.contourArea(cv2.convexHull(pts)) if rc != 0 and pts is not None and len(pts) >= 3 else 0.0
        want = inter / (a[i, 2] * a[i, 3] + b[i, 2] * b[i, 3] - inter)
        assert abs(got[i] - want) < 2e-4, (i, got[i], want)
