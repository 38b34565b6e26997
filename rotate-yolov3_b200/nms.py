"""Rotated NMS behind the reference's Python signatures.

* ``r_nms(dets, threshold)``            <- utils/nms/src/rotate_polygon_nms.cpp:7-16 (pybind module ``r_nms``)
* ``non_max_suppression(prediction, conf_thres, nms_thres)`` <- utils/nms/nms.py:4-69 (the ``use_cuda_nms`` branch)

Everything below the tensor plumbing happens in libryolo.so (ryolo_rnms / ryolo_nms_filter)."""
import ctypes

import torch

from . import _lib


def _workspace(nbytes, device):
    # torch's caching allocator is stream-aware; 256-B alignment is guaranteed (512-B blocks)
    return torch.empty(int(nbytes), dtype=torch.uint8, device=device)


def r_nms(dets, threshold):
    """dets: Tensor[N, 6] float32 CUDA (x, y, w, h, theta, score) -> Tensor[K] int64 of kept ORIGINAL indices,
    ascending (utils/nms/src/rotate_polygon_nms_kernel.cu:380-383).

    Reference behaviours kept: a non-CUDA input raises RuntimeError (CHECK_CUDA, rotate_polygon_nms.cpp:3,8);
    an empty input returns an empty CPU long tensor (:9-10); the call is synchronous."""
    if not isinstance(dets, torch.Tensor) or not dets.is_cuda:
        raise RuntimeError("dets must be a CUDAtensor ")
    if dets.numel() == 0:
        return torch.empty((0,), dtype=torch.long, device="cpu")
    if dets.dim() != 2 or dets.shape[1] < 6:
        raise RuntimeError("r_nms: dets must be [N, 6] (x, y, w, h, theta, score)")
    with torch.cuda.device(dets.device):
        d = dets[:, :6].to(torch.float32).contiguous()
        n = d.shape[0]
        ws_bytes = _lib.lib.ryolo_rnms_workspace_bytes(n)
        ws = _workspace(ws_bytes, d.device)
        keep = torch.empty((n,), dtype=torch.long, device=d.device)
        num = torch.empty((1,), dtype=torch.int32, device=d.device)
        st = _lib.lib.ryolo_rnms(_lib.ptr(d), n, float(threshold), _lib.ptr(keep), _lib.ptr(num), _lib.ptr(ws),
                                 ws_bytes, _lib.stream_ptr(d.device))
        _lib.check(st, "ryolo_rnms")
        k = int(num.item())  # the reference is synchronous too (blocking D2H of the whole mask)
        return keep[:k]


def r_nms_async(dets, threshold):
    """Same computation as r_nms without the host synchronisation: returns (keep [N] int64, num_keep [1] int32) device
    tensors; keep[:num_keep] are the kept original indices, ascending.  dets must be [N>0, 6] float32 CUDA contiguous.
    Used to overlap several images' NMS on different streams (bench.py detect workload)."""
    n = dets.shape[0]
    ws_bytes = _lib.lib.ryolo_rnms_workspace_bytes(n)
    ws = _workspace(ws_bytes, dets.device)
    keep = torch.empty((n,), dtype=torch.long, device=dets.device)
    num = torch.empty((1,), dtype=torch.int32, device=dets.device)
    st = _lib.lib.ryolo_rnms(_lib.ptr(dets), n, float(threshold), _lib.ptr(keep), _lib.ptr(num), _lib.ptr(ws), ws_bytes,
                             _lib.stream_ptr(dets.device))
    _lib.check(st, "ryolo_rnms")
    return keep, num, ws


def nms_filter_async(pred, conf_thres, min_wh, capacity):
    """nms_filter without the host synchronisation: returns (out [capacity, 8] pre-filled with conf = -1, num [1])."""
    p, no = pred.shape
    out = torch.full((capacity, 8), -1.0, dtype=torch.float32, device=pred.device)
    num = torch.empty((1,), dtype=torch.int32, device=pred.device)
    ws_bytes = _lib.lib.ryolo_nms_filter_workspace_bytes(p)
    ws = _workspace(ws_bytes, pred.device)
    st = _lib.lib.ryolo_nms_filter(_lib.ptr(pred), p, no - 6, float(conf_thres), float(min_wh), _lib.ptr(out), capacity,
                                   _lib.ptr(num), _lib.ptr(ws), ws_bytes, _lib.stream_ptr(pred.device))
    _lib.check(st, "ryolo_nms_filter")
    return out, num


def rnms_debug(dets, threshold):
    """Test hook: run r_nms and also return (sorted_boxes [N,6], order [N] int32, mask [N, ceil(N/64)] int64 view)."""
    d = dets[:, :6].to(torch.float32).contiguous()
    n = d.shape[0]
    with torch.cuda.device(d.device):
        ws_bytes = _lib.lib.ryolo_rnms_workspace_bytes(n)
        ws = _workspace(ws_bytes, d.device)
        keep = torch.empty((n,), dtype=torch.long, device=d.device)
        num = torch.empty((1,), dtype=torch.int32, device=d.device)
        st = _lib.lib.ryolo_rnms_full_mask(_lib.ptr(d), n, float(threshold), _lib.ptr(keep), _lib.ptr(num), _lib.ptr(ws),
                                           ws_bytes, _lib.stream_ptr(d.device))
        _lib.check(st, "ryolo_rnms_full_mask")
        k = int(num.item())
        pb, po, pm = ctypes.c_void_p(), ctypes.c_void_p(), ctypes.c_void_p()
        _lib.check(_lib.lib.ryolo_rnms_debug_views(_lib.ptr(ws), n, ctypes.byref(pb), ctypes.byref(po),
                                                   ctypes.byref(pm)), "ryolo_rnms_debug_views")
        base = ws.data_ptr()
        cb = (n + 63) // 64

        def view(p, nbytes, dtype, shape):
            off = p.value - base
            return ws[off:off + nbytes].view(dtype).view(shape).clone()
        boxes = view(pb, n * 6 * 4, torch.float32, (n, 6))
        order = view(po, n * 4, torch.int32, (n,))
        mask = view(pm, n * cb * 8, torch.int64, (n, cb))
        return keep[:k].clone(), boxes, order, mask


def nms_filter(pred, conf_thres, min_wh=2.0):
    """Candidate filter of non_max_suppression for ONE image (utils/nms/nms.py:34-40,55).
    pred [P, 6+nc] float32 CUDA, modified in place (pred[:, 5] *= class_conf).  Returns Tensor[n, 8]
    (x, y, w, h, theta, conf, class_conf, class) in input order."""
    assert pred.is_cuda and pred.dtype == torch.float32 and pred.is_contiguous()
    p, no = pred.shape
    nc = no - 6
    with torch.cuda.device(pred.device):
        out = torch.empty((p, 8), dtype=torch.float32, device=pred.device)
        num = torch.empty((1,), dtype=torch.int32, device=pred.device)
        ws_bytes = _lib.lib.ryolo_nms_filter_workspace_bytes(p)
        ws = _workspace(ws_bytes, pred.device)
        st = _lib.lib.ryolo_nms_filter(_lib.ptr(pred), p, nc, float(conf_thres), float(min_wh), _lib.ptr(out), p,
                                       _lib.ptr(num), _lib.ptr(ws), ws_bytes, _lib.stream_ptr(pred.device))
        _lib.check(st, "ryolo_nms_filter")
        return out[:int(num.item())]


def non_max_suppression(prediction, conf_thres=0.5, nms_thres=0.5):
    """Drop-in for utils/nms/nms.py:4-69: per image, (x, y, w, h, a, object_conf*class_conf, class_conf, class)
    rows of the survivors sorted by confidence, or None.  Mutates prediction[..., 5] like the reference (:35)."""
    min_wh = 2
    output = [None] * len(prediction)
    for image_i, pred in enumerate(prediction):
        if prediction.numel() == 0:
            continue
        if not pred.is_contiguous():  # keep the in-place side effect on the caller's tensor
            raise RuntimeError("non_max_suppression: prediction must be contiguous")
        cand = nms_filter(pred, conf_thres, min_wh)
        if len(cand) == 0:
            continue
        cand = cand[(-cand[:, 5]).argsort()]
        if pred.shape[1] == 7:
            # single-class configs (every cfg the reference ships): class_pred is 0 for every row, so the per-class loop,
            # its unique() (a host sync), the boolean masks and both re-sorts are identities -- r_nms returns ascending
            # indices into the confidence-sorted rows, i.e. the result is already in the reference's final order
            inds = r_nms(cand[:, :6], nms_thres)
            output[image_i] = cand[inds]
            continue
        det_max = []
        for c in cand[:, -1].unique():
            dc = cand[cand[:, -1] == c]
            dc = dc[(-dc[:, 5]).argsort()]
            inds = r_nms(dc[:, :6], nms_thres)
            det_max.append(dc[inds])
        if len(det_max):
            det_max = torch.cat(det_max)
            output[image_i] = det_max[(-det_max[:, 5]).argsort()]
    return output


def r_nms_batched(dets, counts, threshold, limit=None):
    """S independent rotated-NMS problems in one set of launches, no host synchronisation.
    dets [S, cap, 6] float32 CUDA contiguous, counts [S] int32 CUDA (rows used per segment), limit: at most this many
    best-scored boxes per segment enter the NMS.  Returns (keep [S, cap] int64, num_keep [S] int32): segment s keeps rows
    keep[s, :num_keep[s]] (ascending row indices within the segment) -- per segment identical to r_nms(dets[s, :n])."""
    s_, cap = dets.shape[0], dets.shape[1]
    limit = cap if limit is None else int(limit)
    with torch.cuda.device(dets.device):
        ws_bytes = _lib.lib.ryolo_rnms_batched_workspace_bytes(s_, cap)
        ws = _workspace(ws_bytes, dets.device)
        keep = torch.empty((s_, cap), dtype=torch.long, device=dets.device)
        num = torch.empty((s_,), dtype=torch.int32, device=dets.device)
        st = _lib.lib.ryolo_rnms_batched(_lib.ptr(dets), s_, cap, _lib.ptr(counts), limit, float(threshold), _lib.ptr(keep),
                                         _lib.ptr(num), _lib.ptr(ws), ws_bytes, _lib.stream_ptr(dets.device))
        _lib.check(st, "ryolo_rnms_batched")
    return keep, num


def detect_select(io, conf_thres, limit, min_wh=2.0, slack=2048):
    """Candidate selection of a whole decoded batch on the device: per image the `limit` most confident rows that pass
    the reference's filter (utils/nms/nms.py:34-40).  Returns (dets [B, limit+slack, 6], counts [B] int32)."""
    bsz, p, no = io.shape
    cap = int(limit) + int(slack)
    with torch.cuda.device(io.device):
        dets = torch.empty((bsz, cap, 6), dtype=torch.float32, device=io.device)
        counts = torch.empty((bsz,), dtype=torch.int32, device=io.device)
        ws_bytes = _lib.lib.ryolo_detect_select_workspace_bytes(bsz, p)
        ws = _workspace(ws_bytes, io.device)
        st = _lib.lib.ryolo_detect_select(_lib.ptr(io), bsz, p, no - 6, float(conf_thres), float(min_wh), int(limit),
                                          _lib.ptr(dets), cap, _lib.ptr(counts), _lib.ptr(ws), ws_bytes,
                                          _lib.stream_ptr(io.device))
        _lib.check(st, "ryolo_detect_select")
    return dets, counts


def detect_postprocess(io, conf_thres, nms_thres, cap, filter_done_event=None):
    """detect.py:204-213 after the forward, for a whole batch WITHOUT host round trips: device-side candidate selection
    (filter + top-`cap` by confidence; the reference has no cap and relies on trained weights, a random-init network puts
    ~half of the 545 832 proposals above any threshold) and ONE segmented rotated NMS over all images (single-class
    configs: one segment per image).  io [B, P, 6+nc] float32 CUDA contiguous.
    Returns dict(dets [B, cap+slack, 6], counts [B], keep [B, cap+slack] int64, num_keep [B] int32): image b keeps rows
    dets[b, keep[b, :num_keep[b]]]."""
    if not io.is_contiguous():
        io = io.contiguous()
    dets, counts = detect_select(io, conf_thres, cap)
    if filter_done_event is not None:
        filter_done_event.record()
    keep, num = r_nms_batched(dets, counts, nms_thres, limit=cap)
    return {"dets": dets, "counts": counts, "keep": keep, "num_keep": num}
