"""Training-mode execution of ``Darknet`` on libryolo.so: forward with batch-statistics BatchNorm and the full backward
(reference: train.py:268-282 -- ``pred = model(imgs)``; ``loss.backward()`` through nn.Conv2d / BatchNorm2d / PReLU,
shortcut, route, upsample).

Per conv block:  z = conv(x, W) (tcgen05 implicit GEMM, raw, bf16)  ->  batch statistics (bn_stats)  ->
y = prelu(z*scale + shift) [+ shortcut] [2x2 replicated] (bn_act_fwd).  Backward, in reverse:  bn_act_bwd (d-gamma,
d-beta, d-slope, dz in place of z, shortcut gradient)  ->  wgrad (tcgen05 GEMM over pixels)  ->  dgrad (the SAME
implicit-GEMM kernel as forward with the taps mirrored and the weight matrix transposed, accumulating into the
source's gradient buffer through the kernel's residual input).  Stride-2 blocks run their adjoints on the input grid
through a zero-inserted copy of dz.  The first layer (3 channels) goes through an im2col of the image so that it uses
the same kernels.  Activations and activation gradients are bf16 (padded NHWC), parameter gradients fp32.

The whole network is ONE autograd.Function: PyTorch's autograd sees the three head tensors on the outside (so the
loss stays ordinary PyTorch, reference model/loss.py) and receives the parameter gradients in ``parameters()`` order."""
import ctypes

import torch

from . import _lib
from . import layout as L

_EPS = 1e-5


class _Block:
    pass


class TrainPlan:
    def __init__(self, model, batch, height, width, device):
        self.model, self.batch, self.h, self.w, self.device = model, batch, height, width, device
        defs = model.module_defs
        n = len(defs)
        shape = [None] * n
        c, h, w = 3, height, width
        for i, d in enumerate(defs):
            t = d["type"]
            if t == "convolutional":
                s = int(d["stride"])
                c, h, w = int(d["filters"]), (h + s - 1) // s, (w + s - 1) // s
            elif t == "upsample":
                h, w = h * int(d["stride"]), w * int(d["stride"])
            elif t == "route":
                ls = [l if l > 0 else i + l for l in (int(x) for x in d["layers"].split(","))]
                c = sum(shape[l][0] for l in ls)
                h, w = shape[ls[0]][1], shape[ls[0]][2]
            elif t not in ("shortcut", "yolo"):
                raise NotImplementedError("block type %r has no training kernel" % t)
            shape[i] = (c, h, w)
        self.shape = shape

        # activation buffers (y) and their gradient twins; concat groups share one buffer
        target = {}
        views = [None] * n
        gviews = [None] * n

        def alloc_pair(hh, ww, cs):
            return L.alloc_padded(batch, hh, ww, cs, device), L.alloc_padded(batch, hh, ww, cs, device)

        for i, d in enumerate(defs):
            if d["type"] == "route":
                ls = [l if l > 0 else i + l for l in (int(x) for x in d["layers"].split(","))]
                if len(ls) > 1:
                    ctot = sum(shape[l][0] for l in ls)
                    buf, gbuf = alloc_pair(shape[i][1], shape[i][2], L.round_up(ctot, 64))
                    off = 0
                    for l in ls:
                        target[l] = (buf, gbuf, off)
                        off += shape[l][0]
                    views[i] = _mk_view(buf, 0, ctot, shape[i][1], shape[i][2])
                    gviews[i] = _mk_view(gbuf, 0, ctot, shape[i][1], shape[i][2])

        def out_views(layer):
            cc, hh, ww = shape[layer]
            if layer in target:
                buf, gbuf, off = target[layer]
            else:
                buf, gbuf = alloc_pair(hh, ww, L.round_up(cc, 32))
                off = 0
            return _mk_view(buf, off, cc, hh, ww), _mk_view(gbuf, off, cc, hh, ww)

        # first layer input: im2col of the image (27 real channels in a 64-channel padded NHWC buffer)
        self.col = L.alloc_padded(batch, height, width, 64, device)
        blocks = []
        i = 0
        while i < n:
            d = defs[i]
            t = d["type"]
            if t == "convolutional":
                nxt = defs[i + 1]["type"] if i + 1 < n else None
                blk = _Block()
                blk.i = i
                blk.k, blk.stride = int(d["size"]), int(d["stride"])
                seq = model.module_list[i]
                blk.has_bn = hasattr(seq, "BatchNorm2d")
                blk.has_act = hasattr(seq, "activation")
                blk.cout, blk.oh, blk.ow = shape[i]
                if i == 0:
                    if not (blk.k == 3 and blk.stride == 1):
                        raise NotImplementedError("first layer must be 3x3 stride 1")
                    blk.src = _mk_view(self.col, 0, 27, height, width)
                    blk.gsrc = None
                    blk.k_eff = 1                       # runs as a 1x1 conv over the im2col buffer
                else:
                    blk.src, blk.gsrc = views[i - 1], gviews[i - 1]
                    blk.k_eff = blk.k
                blk.s2d = False
                blk.fuse_res = nxt == "shortcut" and i not in model.routes
                blk.fuse_up = nxt == "upsample" and i not in model.routes and int(defs[i + 1]["stride"]) == 2
                blk.is_head = nxt == "yolo"
                blk.mat = i + 1 if (blk.fuse_res or blk.fuse_up) else i
                bn_ = 256 if blk.cout > 128 else (128 if blk.cout > 64 else 64)
                blk.cout_pad = L.round_up(blk.cout, bn_)     # GEMM tile extent (weights / dW)
                blk.zcs = L.round_up(blk.cout, 32)           # channel stride of z / dz (TMA zero-fills up to the tile)
                blk.cin = blk.src.c
                blk.cin_pad = L.round_up(blk.cin, 64)
                if blk.is_head:
                    blk.out = torch.empty((batch, blk.cout, blk.oh, blk.ow), dtype=torch.float32, device=device)
                    blk.dz = L.alloc_padded(batch, blk.oh, blk.ow, blk.zcs, device)   # gradient of the head output
                    blk.res = blk.gres = None
                else:
                    blk.z = L.alloc_padded(batch, blk.oh, blk.ow, blk.zcs, device)     # raw conv output / dz
                    blk.y, blk.gy = out_views(blk.mat)
                    views[blk.mat], gviews[blk.mat] = blk.y, blk.gy
                    if blk.fuse_res:
                        frm = int(defs[i + 1]["from"])
                        ridx = i + 1 + frm if frm < 0 else frm
                        blk.res, blk.gres = views[ridx], gviews[ridx]
                    else:
                        blk.res = blk.gres = None
                    blk.sums = torch.zeros(2 * blk.cout + 1, dtype=torch.float32, device=device)      # + the CTA ticket of the fused finalize
                    blk.mean, blk.invstd, blk.scale, blk.shift = (torch.zeros(blk.cout, dtype=torch.float32, device=device)
                                                                  for _ in range(4))
                blk.s2d = (i > 0 and blk.stride == 2 and blk.k == 3 and blk.src.h % 2 == 0 and blk.src.w % 2 == 0
                           and blk.src.c % 8 == 0)
                if blk.s2d:
                    # 3x3/stride-2 block on the space-to-depth copy of its input: 2x2 taps, 4C channels, stride 1
                    blk.xs_cs = L.round_up(4 * blk.src.c, 64)
                    blk.xs = L.alloc_padded(batch, blk.src.h // 2, blk.src.w // 2, blk.xs_cs, device)
                    blk.dxs = L.alloc_padded(batch, blk.src.h // 2, blk.src.w // 2, blk.xs_cs, device)
                    blk.k_eff = 2
                    blk.cin_eff, blk.cin_pad = 4 * blk.src.c, blk.xs_cs
                    blk.fdesc = L.make_desc(batch, blk.src.h // 2, blk.src.w // 2, blk.cin_eff, blk.xs_cs, blk.cout,
                                            0 if blk.is_head else blk.zcs, 2, 1, False, 0.0, False, 0, False, blk.is_head)
                    blk.ddesc = L.make_desc(batch, blk.src.h // 2, blk.src.w // 2, blk.cout, blk.zcs, blk.cin_eff,
                                            blk.xs_cs, -2, 1, False, 0.0, False, 0, False, False)
                else:
                    blk.cin_eff = blk.cin
                    if blk.stride == 2:
                        blk.dz_up = L.alloc_padded(batch, blk.src.h, blk.src.w, blk.zcs, device)  # zero-inserted dz
                    # forward descriptor: raw conv (no bias / activation); heads keep their bias and write fp32 NCHW
                    blk.fdesc = L.make_desc(batch, blk.src.h, blk.src.w, blk.cin, blk.src.cs, blk.cout,
                                            0 if blk.is_head else blk.zcs, blk.k_eff, blk.stride, False, 0.0, False, 0,
                                            False, blk.is_head)
                    if i > 0:
                        # dgrad: stride-1 conv of dz (at the INPUT resolution) with mirrored taps / transposed weights
                        blk.ddesc = L.make_desc(batch, blk.src.h, blk.src.w, blk.cout, blk.zcs, blk.cin, blk.gsrc.cs,
                                                blk.k, 1, False, 0.0, False, blk.gsrc.cs, False, False)
                blk.dw_shape = (blk.k_eff * blk.k_eff, blk.cout_pad, blk.cin_pad)
                blk.pw = torch.empty(_lib.lib.ryolo_conv_packed_weight_bytes(ctypes.byref(blk.fdesc)), dtype=torch.uint8,
                                     device=device)
                if i > 0:
                    blk.pwd = torch.empty(_lib.lib.ryolo_conv_packed_weight_bytes(ctypes.byref(blk.ddesc)), dtype=torch.uint8,
                                          device=device)
                blocks.append(blk)
                i += 2 if (blk.fuse_res or blk.fuse_up) else 1
                continue
            if t == "route":
                ls = [l if l > 0 else i + l for l in (int(x) for x in d["layers"].split(","))]
                if len(ls) == 1:
                    views[i], gviews[i] = views[ls[0]], gviews[ls[0]]
            elif t in ("shortcut", "upsample"):
                raise NotImplementedError("%s at block %d does not follow a convolution it can be fused into" % (t, i))
            i += 1
        self.blocks = blocks
        # a block whose output feeds a space-to-depth (3x3 / stride 2) block writes the s2d copy itself (bn_act_fwd_s2d)
        for c_ in blocks:
            c_.xs_from_producer = False
        for c_ in blocks:
            if c_.s2d:
                prod = [p_ for p_ in blocks if not p_.is_head and p_.y is c_.src and not p_.fuse_up]
                if len(prod) == 1:
                    prod[0].xs_out = c_.xs
                    c_.xs_from_producer = True
        self.zero_bias = torch.zeros(2048, dtype=torch.float32, device=device)
        # static accumulate flags for the backward writers (first writer of a gradient range overwrites)
        written = {}

        def claim(view):
            key = id(view.buf)
            lo, hi = view.ch_off, view.ch_off + view.c
            covered = any(a <= lo and hi <= b for a, b in written.get(key, []))
            written.setdefault(key, []).append((lo, hi))
            return covered
        for blk in reversed(blocks):
            blk.gres_acc = claim(blk.gres) if blk.gres is not None else False
            blk.gsrc_acc = claim(blk.gsrc) if blk.gsrc is not None else False
        # A producer whose output gradient has ONE writer -- the depth_to_space of its space-to-depth consumer -- reads that
        # gradient straight from the consumer's dgrad output (bn_act_bwd mode 2): no d2s pass, no gy buffer traffic.
        for c_ in blocks:
            c_.dxs_direct = False
        for c_ in blocks:
            if not c_.s2d or c_.gsrc is None:
                continue
            key = id(c_.gsrc.buf)
            lo, hi = c_.gsrc.ch_off, c_.gsrc.ch_off + c_.gsrc.c
            writers = [r for r in written.get(key, []) if not (r[1] <= lo or hi <= r[0])]
            prod = [p_ for p_ in blocks if not p_.is_head and p_.gy is c_.gsrc and not p_.fuse_up]
            if len(writers) == 1 and len(prod) == 1:
                prod[0].dy_s2d = c_
                c_.dxs_direct = True
        self._build_arena()
        # split-K partial sums of every layer's weight gradient in ONE buffer: one fill per step instead of 75
        off = 0
        for blk in blocks:
            blk.dw_off = off
            off += L.round_up(blk.dw_shape[0] * blk.dw_shape[1] * blk.dw_shape[2], 64)
        self.dw_arena = torch.zeros(off, dtype=torch.float32, device=device)
        for blk in blocks:
            n_ = blk.dw_shape[0] * blk.dw_shape[1] * blk.dw_shape[2]
            blk.dw = self.dw_arena[blk.dw_off:blk.dw_off + n_].view(blk.dw_shape)
        self._key_now = None
        self._multi = None           # job tables of the multi-tensor pack / unpack launches (keyed on parameter storage)
        self.gen = 0                 # forward generation: backward must see the activations of ITS forward
        self.consumed = True
        self.graphs = None           # CUDA graphs of forward / backward segments (model.use_cuda_graph)
        self.replayed_kernels = 0    # kernels of this library executed through CUDA-graph replays (bench.py gpu_launches)
        self.time_comm = False       # bench.py: record events around the wait for the all-reduces (exposed time)
        self._comm_ev = None

    def _build_arena(self):
        """All parameter gradients live in ONE flat fp32 arena, laid out in the order backward produces them (last block
        first): per block [conv weight | BN (d beta, d gamma) + PReLU d slope] (heads: [weight | bias]).  The kernels write
        straight into it; contiguous ranges of it are the all-reduce buckets (parallel.GradBuckets) and autograd receives
        views of a fresh scaled copy (never of plan-owned memory: AccumulateGrad may keep what it is given)."""
        m = self.model
        off = 0
        self.param_slots = {}        # (block index, "Module.param") -> (offset, numel, shape)
        bounds = []
        for blk in reversed(self.blocks):
            seq = m.module_list[blk.i]
            wshape = tuple(seq.Conv2d.weight.shape)
            n = seq.Conv2d.weight.numel()
            blk.gw_off = off
            self.param_slots[(blk.i, "Conv2d.weight")] = (off, n, wshape)
            off += L.round_up(n, 4)
            if blk.is_head:
                if seq.Conv2d.bias is not None:
                    blk.gb_off = off
                    self.param_slots[(blk.i, "Conv2d.bias")] = (off, blk.cout, (blk.cout,))
                    off += L.round_up(blk.cout, 4)
            else:
                blk.bs_off = off
                self.param_slots[(blk.i, "BatchNorm2d.bias")] = (off, blk.cout, (blk.cout,))
                self.param_slots[(blk.i, "BatchNorm2d.weight")] = (off + blk.cout, blk.cout, (blk.cout,))
                if blk.has_act:
                    self.param_slots[(blk.i, "activation.weight")] = (off + 2 * blk.cout, 1, (1,))
                off += L.round_up(2 * blk.cout + 1, 4)
            bounds.append(off)
            blk.arena_end = off
        self.garena = torch.zeros(off, dtype=torch.float32, device=self.device)
        for blk in self.blocks:
            seq = m.module_list[blk.i]
            n = seq.Conv2d.weight.numel()
            blk.gw = self.garena[blk.gw_off:blk.gw_off + n].view(tuple(seq.Conv2d.weight.shape))
            if not blk.is_head:
                blk.bsums = self.garena[blk.bs_off:blk.bs_off + 2 * blk.cout + 1]
        ddp = getattr(m, "_ddp", None)
        from .parallel import GradBuckets
        self.buckets = GradBuckets(self.garena, bounds, bucket_bytes=ddp["bucket_bytes"] if ddp else (1 << 62),
                                   group=ddp["group"] if ddp else None) if ddp else None
        # backward segments = runs of blocks whose gradients complete one bucket (one segment without DDP)
        segs, cur = [], []
        ends = set(hi for _, hi in self.buckets.buckets) if self.buckets else set()
        for blk in reversed(self.blocks):
            cur.append(blk)
            if blk.arena_end in ends:
                segs.append(cur)
                cur = []
        if cur:
            segs.append(cur)
        self.segments = segs
        self.names = [nm for nm, _ in m.named_parameters()]

    # ------------------------------------------------------------------------------------------------------
    def _slopes(self):
        # PReLU slopes stay on the device (kernels read nn.PReLU.weight through slope_dev): no host synchronisation
        m = self.model
        for b in self.blocks:
            if b.has_act:
                wgt = m.module_list[b.i].activation.weight
                if wgt.numel() != 1 or wgt.dtype != torch.float32:
                    raise NotImplementedError("PReLU with per-channel or non-fp32 slope")
                b.slope_dev = wgt.data_ptr()
            else:
                b.slope_dev = None
            b.slope = 1.0

    def _param_key(self):
        """storage identity of everything the captured graphs point at"""
        return tuple(p.data_ptr() for p in self.model.parameters()) + tuple(b.data_ptr() for b in self.model.buffers())

    # ---- multi-tensor pack / unpack (csrc/multi.cu): job tables in device memory, rebuilt when parameter storage moves ----
    def _job_table(self, jobs, cls):
        arr = (cls * len(jobs))(*jobs)
        raw = torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8).to(self.device)
        sizes = []
        for j in jobs:
            if cls is _lib.PackJob:
                sizes.append(j.ks * j.ks * j.cout_pad * j.cin_pad)
            else:
                sizes.append(j.cout * j.cin * j.ks * j.ks)
        prefix = torch.tensor([0] + list(torch.tensor(sizes, dtype=torch.int64).cumsum(0).tolist()), dtype=torch.int64)
        return raw, prefix.to(self.device), len(jobs), int(prefix[-1])

    def _multi_tables(self):
        key = self._key_now if self._key_now is not None else self._param_key()
        if self._multi is not None and self._multi["key"] == key:
            return self._multi
        m = self.model

        def pads(cout, cin):
            bn_ = 256 if cout > 128 else (128 if cout > 64 else 64)
            return L.round_up(cout, bn_), L.round_up(cin, 64)
        pack = []
        for blk in self.blocks:
            w = m.module_list[blk.i].Conv2d.weight
            if blk.i == 0:
                cout, cin, ks, mode = blk.cout, 27, 1, 0                       # im2col form: [cout, 27, 1, 1]
            elif blk.s2d:
                cout, cin, ks, mode = blk.cout, blk.cin_eff, 2, 2
            else:
                cout, cin, ks, mode = blk.cout, blk.cin, blk.k_eff, 0
            cp, kp = pads(cout, cin)
            assert blk.pw.numel() == ks * ks * cp * kp * 2, (blk.i, blk.pw.numel(), ks, cp, kp)
            pack.append(_lib.PackJob(w.data_ptr(), blk.pw.data_ptr(), cout, cin, ks, cp, kp, mode))
            if blk.i > 0:                                                        # dgrad operand: mirrored taps, transposed
                if blk.s2d:
                    cout_d, cin_d, ks_d, mode_d = blk.cin_eff, blk.cout, 2, 3
                else:
                    cout_d, cin_d, ks_d, mode_d = blk.cin, blk.cout, blk.k, 1
                cpd, kpd = pads(cout_d, cin_d)
                assert blk.pwd.numel() == ks_d * ks_d * cpd * kpd * 2, (blk.i, blk.pwd.numel(), ks_d, cpd, kpd)
                pack.append(_lib.PackJob(w.data_ptr(), blk.pwd.data_ptr(), cout_d, cin_d, ks_d, cpd, kpd, mode_d))
        unpack = []
        for seg in self.segments:
            jobs = []
            for blk in seg:
                if blk.s2d:
                    jobs.append(_lib.UnpackJob(blk.dw.data_ptr(), blk.gw.data_ptr(), blk.cout_pad, blk.cin_pad, 2, blk.cout,
                                               blk.src.c, 3))
                else:
                    jobs.append(_lib.UnpackJob(blk.dw.data_ptr(), blk.gw.data_ptr(), blk.cout_pad, blk.cin_pad, 0, blk.cout,
                                               blk.cin, blk.k_eff))
            unpack.append(self._job_table(jobs, _lib.UnpackJob))
        self._multi = {"key": key, "pack": self._job_table(pack, _lib.PackJob), "unpack": unpack}
        return self._multi

    def _pack_all(self):
        """every forward AND dgrad weight operand of the step in one launch (the weights are constant within a step)"""
        for b in self.blocks:
            if self.model.module_list[b.i].Conv2d.weight.dtype != torch.float32:
                raise RuntimeError("fp32 parameters required")
        raw, prefix, n, total = self._multi_tables()["pack"]
        _lib.check(_lib.lib.ryolo_conv_pack_weights_multi(_lib.ptr(raw), n, _lib.ptr(prefix), total, _lib.stream_ptr(self.device)),
                   "pack_weights_multi")

    def _backward_segment(self, si):
        if si == 0:
            self.dw_arena.zero_()
        for blk in self.segments[si]:
            self._backward_block(blk)
        raw, prefix, n, total = self._multi_tables()["unpack"][si]
        _lib.check(_lib.lib.ryolo_conv_unpack_wgrad_multi(_lib.ptr(raw), n, _lib.ptr(prefix), total,
                                                          _lib.stream_ptr(self.device)), "unpack_wgrad_multi")

    def _forward_body(self, x):
        m = self.model
        if any(p.dtype != torch.float32 for p in m.parameters()):
            # .half() models are supported in eval mode (weights are re-packed to bf16 anyway); the fused training step
            # reads and updates fp32 master parameters in place, like apex O1 keeps them (train.py:165-166)
            raise RuntimeError("Darknet training step needs fp32 parameters (call .float(); .half() is an inference option)")
        lib = _lib.lib
        st = _lib.stream_ptr(self.device)
        _lib.check(lib.ryolo_im2col_first(_lib.ptr(x), self.batch, self.h, self.w, _lib.ptr(self.col), st), "im2col")
        n_per_pixel = float(self.batch)
        self._slopes()
        self._pack_all()
        bn_counters = []
        for blk in self.blocks:
            seq = m.module_list[blk.i]
            w = seq.Conv2d.weight.detach()
            if blk.i == 0:
                w = w.reshape(blk.cout, 27, 1, 1)
            x_ptr = blk.src.ptr
            if blk.s2d:
                if not blk.xs_from_producer:
                    _lib.check(lib.ryolo_space_to_depth(ctypes.c_void_p(blk.src.ptr), blk.src.cs, self.batch, blk.src.h,
                                                        blk.src.w, blk.src.c, _lib.ptr(blk.xs), blk.xs_cs, st), "s2d")
                x_ptr = blk.xs.data_ptr()
            if blk.is_head:
                if getattr(blk, "bias_pad", None) is None:
                    blk.bias_pad = L.padded_bias(blk.fdesc, torch.zeros(blk.cout, device=self.device))
                if seq.Conv2d.bias is not None:
                    blk.bias_pad[:blk.cout].copy_(seq.Conv2d.bias.detach())
                _lib.check(lib.ryolo_conv_bn_act_fwd(ctypes.byref(blk.fdesc), ctypes.c_void_p(x_ptr), _lib.ptr(blk.pw),
                                                     _lib.ptr(blk.bias_pad), None, _lib.ptr(blk.out), None, 0, st), "conv head")
                continue
            _lib.check(lib.ryolo_conv_bn_act_fwd(ctypes.byref(blk.fdesc), ctypes.c_void_p(x_ptr), _lib.ptr(blk.pw),
                                                 _lib.ptr(self.zero_bias), None, _lib.ptr(blk.z), None, 0, st), "conv")
            cnt = n_per_pixel * blk.oh * blk.ow
            if blk.has_bn:
                bn = seq.BatchNorm2d
                _lib.check(lib.ryolo_bn_stats_finalize(_lib.ptr(blk.z), blk.zcs, self.batch, blk.oh, blk.ow, blk.cout,
                                                       _lib.ptr(blk.sums), bn.eps, bn.momentum, _lib.ptr(bn.weight),
                                                       _lib.ptr(bn.bias), _lib.ptr(blk.mean), _lib.ptr(blk.invstd),
                                                       _lib.ptr(blk.scale), _lib.ptr(blk.shift), _lib.ptr(bn.running_mean),
                                                       _lib.ptr(bn.running_var), st), "bn_stats_finalize")
                bn_counters.append(bn.num_batches_tracked)
            else:
                raise NotImplementedError("conv block without BatchNorm that is not a YOLO head")
            xs_out = getattr(blk, "xs_out", None)
            if xs_out is not None:
                _lib.check(lib.ryolo_bn_act_fwd_s2d(_lib.ptr(blk.z), blk.zcs, self.batch, blk.oh, blk.ow, blk.cout,
                                                    _lib.ptr(blk.scale), _lib.ptr(blk.shift), blk.slope, int(blk.has_act),
                                                    ctypes.c_void_p(blk.res.ptr) if blk.res is not None else None,
                                                    blk.res.cs if blk.res is not None else 0, ctypes.c_void_p(blk.y.ptr),
                                                    blk.y.cs, ctypes.c_void_p(blk.slope_dev), _lib.ptr(xs_out),
                                                    xs_out.shape[-1], st), "bn_act_fwd_s2d")
                continue
            _lib.check(lib.ryolo_bn_act_fwd(_lib.ptr(blk.z), blk.zcs, self.batch, blk.oh, blk.ow, blk.cout,
                                            _lib.ptr(blk.scale), _lib.ptr(blk.shift), blk.slope, int(blk.has_act),
                                            ctypes.c_void_p(blk.res.ptr) if blk.res is not None else None,
                                            blk.res.cs if blk.res is not None else 0, ctypes.c_void_p(blk.y.ptr), blk.y.cs,
                                            int(blk.fuse_up), ctypes.c_void_p(blk.slope_dev), st), "bn_act_fwd")
        if bn_counters:
            with torch.no_grad():
                torch._foreach_add_(bn_counters, 1)
        # the fp32 NCHW buffers the head convolutions wrote; DarknetTrainFn's caller hands out permuted VIEWS of them
        # ([B, na, ny, nx, no] like model/models.py:190-192, without the .contiguous() copy of 0.5 GB per step)
        return [blk.out for blk in self.heads]

    def forward(self, x):
        m = self.model
        self.heads = [b for b in self.blocks if b.is_head]
        for blk, yi in zip(self.heads, m.yolo_layers):
            layer = m.module_list[yi]
            if (layer.nx, layer.ny) != (blk.ow, blk.oh):
                layer.create_grids((self.h, self.w), (blk.ow, blk.oh), self.device, torch.float32)
        self._key_now = self._param_key()      # once per step: storage identity of parameters / buffers
        if getattr(m, "use_cuda_graph", False):
            outs = self._forward_graph(x)
        else:
            outs = self._forward_body(x)
        self.gen += 1
        self.consumed = False
        return outs

    # ---- CUDA-graph form (opt-in, model.use_cuda_graph): the ~1100 launches of a forward and the ~1300 of a backward are
    # replayed as a handful of graphs -- what makes the step launch-bound at small per-GPU batches (8 GPUs: 8 images per
    # rank) is the host, not the GPU.  Inputs are copied into static buffers; outputs are static buffers that the next
    # step overwrites (standard CUDA-graph semantics).
    def _forward_graph(self, x):
        key = self._key_now
        if self.graphs is None or self.graphs["key"] != key:
            self.graphs = {"key": key, "pool": None}
            self.x_static = x.clone()
            eager = self._forward_body(self.x_static)         # this call's forward, eagerly (also the capture warm-up)
            torch.cuda.synchronize(self.device)
            g = torch.cuda.CUDAGraph()
            c0 = _lib.lib.ryolo_launch_count()
            with torch.cuda.graph(g):
                outs = self._forward_body(self.x_static)      # captured, not executed
            self.graphs["pool"] = g.pool()
            self.graphs["fwd"] = (g, outs)
            self.graphs["fwd_kernels"] = int(_lib.lib.ryolo_launch_count() - c0)   # this library's kernels in the graph
            return eager
        g, outs = self.graphs["fwd"]
        self.x_static.copy_(x)
        g.replay()
        self.replayed_kernels += self.graphs["fwd_kernels"]
        return outs

    # ------------------------------------------------------------------------------------------------------
    def _head_grads(self, grads):
        """autograd's head cotangents, fp32 NCHW [B, na*no, ny, nx] (the layout of the head buffers; contiguous without a
        copy when the consumer differentiated the permuted view in place, as loss._FusedLoss does) -> bias gradients +
        padded-NHWC bf16 dz of the head convs"""
        lib = _lib.lib
        st = _lib.stream_ptr(self.device)
        for blk, g in zip(self.heads, grads):
            g = g.contiguous().float()
            gb = self.garena[blk.gb_off:blk.gb_off + blk.cout] if (blk.i, "Conv2d.bias") in self.param_slots else None
            if gb is not None:       # the bias gradient = per-channel sums of g, taken on the way by the same kernel
                gb.zero_()
            _lib.check(lib.ryolo_head_grad_nchw_to_padded(_lib.ptr(g), self.batch, blk.cout, blk.oh, blk.ow, _lib.ptr(blk.dz),
                                                          blk.zcs, _lib.ptr(gb) if gb is not None else None, st),
                       "head_grad_nchw_to_padded")

    def _backward_block(self, blk):
        m = self.model
        lib = _lib.lib
        st = _lib.stream_ptr(self.device)
        seq = m.module_list[blk.i]
        if blk.is_head:
            dz = blk.dz
        else:
            src_c = getattr(blk, "dy_s2d", None)
            dy_ptr, dy_cs, dy_mode = (blk.gy.ptr, blk.gy.cs, int(blk.fuse_up)) if src_c is None else \
                (src_c.dxs.data_ptr(), src_c.xs_cs, 2)
            _lib.check(lib.ryolo_bn_act_bwd(ctypes.c_void_p(dy_ptr), dy_cs, dy_mode, _lib.ptr(blk.z),
                                            blk.zcs, self.batch, blk.oh, blk.ow, blk.cout, _lib.ptr(blk.scale),
                                            _lib.ptr(blk.shift), _lib.ptr(blk.mean), _lib.ptr(blk.invstd), blk.slope,
                                            int(blk.has_act), 1, _lib.ptr(blk.bsums),
                                            ctypes.c_void_p(blk.gres.ptr) if blk.gres is not None else None,
                                            blk.gres.cs if blk.gres is not None else 0, int(blk.gres_acc),
                                            ctypes.c_void_p(blk.slope_dev), st),
                       "bn_act_bwd")
            dz = blk.z
        if blk.s2d:
            _lib.check(lib.ryolo_conv_wgrad(_lib.ptr(dz), blk.zcs, blk.cout_pad, _lib.ptr(blk.xs), blk.xs_cs,
                                            blk.cin_pad, self.batch, blk.src.h // 2, blk.src.w // 2, 2, _lib.ptr(blk.dw),
                                            st), "wgrad s2d")
            _lib.check(lib.ryolo_conv_bn_act_fwd(ctypes.byref(blk.ddesc), _lib.ptr(dz), _lib.ptr(blk.pwd),
                                                 _lib.ptr(self.zero_bias), None, _lib.ptr(blk.dxs), None, 0, st), "dgrad s2d")
            if not blk.dxs_direct:
                _lib.check(lib.ryolo_depth_to_space(_lib.ptr(blk.dxs), blk.xs_cs, self.batch, blk.src.h, blk.src.w, blk.src.c,
                                                    ctypes.c_void_p(blk.gsrc.ptr), blk.gsrc.cs, int(blk.gsrc_acc), st), "d2s")
            return
        if blk.stride == 2:
            _lib.check(lib.ryolo_zero_insert2x(_lib.ptr(dz), blk.zcs, self.batch, blk.oh, blk.ow, blk.zcs,
                                               _lib.ptr(blk.dz_up), blk.zcs, blk.src.h, blk.src.w, st), "zero_insert")
            dz = blk.dz_up
        _lib.check(lib.ryolo_conv_wgrad(_lib.ptr(dz), blk.zcs, blk.cout_pad, ctypes.c_void_p(blk.src.ptr), blk.src.cs,
                                        blk.cin_pad, self.batch, blk.src.h, blk.src.w, blk.k_eff, _lib.ptr(blk.dw), st),
                   "wgrad")
        if blk.i > 0:
            gp = blk.gsrc.ptr
            blk.ddesc.has_residual = int(blk.gsrc_acc)
            _lib.check(lib.ryolo_conv_bn_act_fwd(ctypes.byref(blk.ddesc), _lib.ptr(dz), _lib.ptr(blk.pwd),
                                                 _lib.ptr(self.zero_bias), ctypes.c_void_p(gp) if blk.gsrc_acc else None,
                                                 ctypes.c_void_p(gp), None, 0, st), "dgrad")

    def backward(self, grads):
        m = self.model
        overlap = self.buckets is not None and self.buckets.world > 1
        if overlap:
            # leave a few SMs to the NCCL all-reduce kernels that run next to the backward GEMMs (persistent kernels with a
            # static tile schedule would otherwise run a second wave on the SMs NCCL occupies)
            _lib.lib.ryolo_set_reserved_sms(int(getattr(m, "_ddp", {}).get("reserved_sms", 0)))
        try:
            return self._backward(grads)
        finally:
            if overlap:
                _lib.lib.ryolo_set_reserved_sms(0)

    def _backward(self, grads):
        m = self.model
        self._head_grads(grads)
        use_graph = getattr(m, "use_cuda_graph", False) and self.graphs is not None and "fwd" in self.graphs
        if use_graph and "bwd" not in self.graphs:
            # first backward after capture of the forward: run eagerly once (warm-up), then capture every segment
            for si, seg in enumerate(self.segments):
                self._backward_segment(si)
                if self.buckets:
                    self.buckets.ready(seg[-1].arena_end)
            inv = self.buckets.finish() if self.buckets else 1.0
            out_flat = self.garena * inv
            torch.cuda.synchronize(self.device)
            gs, gk = [], []
            for si in range(len(self.segments)):
                g = torch.cuda.CUDAGraph()
                c0 = _lib.lib.ryolo_launch_count()
                with torch.cuda.graph(g, pool=self.graphs["pool"]):
                    self._backward_segment(si)
                gs.append(g)
                gk.append(int(_lib.lib.ryolo_launch_count() - c0))
            self.graphs["bwd_kernels"] = gk
            self.graphs["bwd"] = gs
            # the capture passes did not execute; the eager pass above produced this step's gradients
        else:
            for si, seg in enumerate(self.segments):
                if use_graph:
                    self.graphs["bwd"][si].replay()
                    self.replayed_kernels += self.graphs["bwd_kernels"][si]
                else:
                    self._backward_segment(si)
                if self.buckets:
                    self.buckets.ready(seg[-1].arena_end)     # NCCL all-reduce of the finished bucket on the side stream
            if self.time_comm:
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
            inv = self.buckets.finish() if self.buckets else 1.0
            if self.time_comm:
                e1.record()
                self._comm_ev = (e0, e1)
            out_flat = self.garena * inv                       # fresh tensor: autograd may keep what it is given
        self.consumed = True
        out = []
        for name in self.names:
            parts = name.split(".")            # module_list.{i}.{Module}.{param}
            slot = self.param_slots.get((int(parts[1]), parts[2] + "." + parts[3]))
            out.append(out_flat[slot[0]:slot[0] + slot[1]].view(slot[2]) if slot is not None else None)
        return out


def _comm_wait_ms(self):
    """time the main stream spent waiting for the gradient all-reduces AFTER the last backward kernel was enqueued = the
    exposed (non-overlapped) part of the collective; None when not measured"""
    if self._comm_ev is None:
        return None
    self._comm_ev[1].synchronize()
    return self._comm_ev[0].elapsed_time(self._comm_ev[1])


TrainPlan.last_comm_wait_ms = _comm_wait_ms


def _mk_view(buf, ch_off, c, h, w):
    from .models import _View
    return _View(buf, ch_off, c, h, w)


class DarknetTrainFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, plan, x, *params):
        ctx.plan = plan
        outs = plan.forward(x)
        ctx.gen = plan.gen
        return tuple(outs)

    @staticmethod
    def backward(ctx, *grads):
        plan = ctx.plan
        if plan.gen != ctx.gen or plan.consumed:
            # activations live in the plan (not per call): a second forward overwrote them, or a backward already turned
            # z into dz in place.  Stock autograd would handle or reject these cases; silently wrong gradients are not an option.
            raise RuntimeError("Darknet training plan: the activations of this forward were overwritten by a later forward "
                               "or already consumed by a backward (retain_graph / two forwards before one backward are "
                               "not supported by the fused training path)")
        gl = []
        for g, blk, yi in zip(grads, plan.heads, plan.model.yolo_layers):
            if g is None:
                g = torch.zeros((plan.batch, blk.cout, blk.oh, blk.ow), dtype=torch.float32, device=plan.device)
            gl.append(g)
        pg = plan.backward(gl)
        return (None, None) + tuple(pg)
