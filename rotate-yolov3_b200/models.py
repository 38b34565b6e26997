"""``Darknet(cfg, hyp, arc).forward`` behind the reference's signature (model/models.py:230-313), executed by
libryolo.so: every Conv2d -> BatchNorm2d -> PReLU block is one tcgen05 implicit-GEMM launch with BN folded, the
shortcut add / nearest-x2 upsample / channel concat of the graph fused into the producing conv's epilogue (residual
read, 2x2 replicated store, store at a channel offset of the shared concat buffer), and the YOLO heads decoded by one
fused kernel each.

What is kept from the reference so that train.py / test.py / detect.py see the same object:
  * ``module_defs`` / ``module_list`` / ``routes`` / ``yolo_layers`` / ``hyp`` / ``version`` / ``seen``;
  * parameter names: ``module_list.{i}.Conv2d.weight``, ``.BatchNorm2d.{weight,bias,running_mean,running_var,
    num_batches_tracked}``, ``.activation.weight`` (checkpoints and the optimizer split key on them);
  * YOLO layers expose ``ng, anchor_vec, anchor_wh, stride, nx, ny, na, nc`` (read by model/loss.py:172-174,313);
  * eval forward returns ``(io [B, sum(na*ny*nx), nc+6], (p0, p1, p2))`` with p_i = [B, na, ny, nx, nc+6].

Eval mode: inference path above.  Training mode: ``forward`` returns the list of raw head tensors like the reference
(models.py:192-194) and is differentiable -- forward with batch-statistics BatchNorm and the whole backward run on the
same library (train_path.py).  There is no PyTorch fallback in either mode."""
import ctypes

import numpy as np
import torch
import torch.nn as nn

from . import _lib
from . import layout as L
from .parse_config import parse_model_cfg


class SELayer(nn.Module):
    """reference model/models.py:16-31 (same parameter names: fc.0.weight [c/r, c], fc.2.weight [c, c/r]); executed by
    ryolo_se_block inside the eval plan, this forward exists only for module-tree parity."""

    def __init__(self, channel, reduction=16):
        super().__init__()
        self.avg_pool = nn.AdaptiveAvgPool2d(1)
        self.fc = nn.Sequential(nn.Linear(channel, channel // reduction, bias=False), nn.ReLU(inplace=True),
                                nn.Linear(channel // reduction, channel, bias=False), nn.Sigmoid())

    def forward(self, x):
        raise RuntimeError("SELayer runs inside Darknet's fused plan (libryolo.so), not as a stand-alone module")


class YOLOLayer(nn.Module):
    """reference model/models.py:170-227; decode runs in ryolo_yolo_decode."""

    def __init__(self, anchors, nc, yolo_index, arc, hyp):
        super().__init__()
        self.anchors = torch.Tensor(np.asarray(anchors, dtype=np.float64))
        self.na = len(anchors)
        self.nc = nc
        self.nx = 0
        self.ny = 0
        self.arc = arc
        self.hyp = hyp
        self.yolo_index = yolo_index

    def create_grids(self, img_size, ng, device, dtype=torch.float32):
        """reference model/model_utils.py:16-35 (without its in-place division bug on CPU: anchors are not
        modified, anchor_vec is a fresh tensor)."""
        nx, ny = ng
        self.img_size = max(img_size)
        self.stride = self.img_size / max(ng)
        yv, xv = torch.meshgrid([torch.arange(ny), torch.arange(nx)], indexing="ij")
        self.grid_xy = torch.stack((xv, yv), 2).to(device).type(dtype).view((1, 1, ny, nx, 2))
        self.anchor_vec = self.anchors.clone().to(device)
        self.anchor_vec[:, :2] /= self.stride
        self.anchor_wh = self.anchor_vec.view(1, self.na, 1, 1, 3).to(device).type(dtype)
        self.ng = torch.Tensor(ng).to(device)
        self.nx = nx
        self.ny = ny
        self._anchors_dev = self.anchors.to(device=device, dtype=torch.float32).contiguous()

    def decode_into(self, p, img_size, io, rows_total, row_offset):
        """p: [B, na*(nc+6), ny, nx] fp32 NCHW head output -> writes rows of io, returns the permuted raw view."""
        bs, ny, nx = p.shape[0], p.shape[-2], p.shape[-1]
        if (self.nx, self.ny) != (nx, ny) or getattr(self, "_anchors_dev", None) is None or \
                self._anchors_dev.device != p.device:
            self.create_grids(img_size, (nx, ny), p.device, p.dtype)
        # the raw head is handed out as the [B, na, ny, nx, no] VIEW of the NCHW buffer (model/models.py:190-192 without the
        # .contiguous() copy: a third of this kernel's traffic), like the training path
        st = _lib.lib.ryolo_yolo_decode(_lib.ptr(p), bs, self.na, self.nc, ny, nx, _lib.ptr(self._anchors_dev),
                                        float(self.stride), float(self.hyp["context_factor"]),
                                        1 if "default" in self.arc else 0, _lib.ptr(io), rows_total, row_offset,
                                        None, _lib.stream_ptr(p.device))
        _lib.check(st, "ryolo_yolo_decode")
        return p.view(bs, self.na, self.nc + 6, ny, nx).permute(0, 1, 3, 4, 2)

    def forward(self, p, img_size, var=None):
        if self.training:
            bs, ny, nx = p.shape[0], p.shape[-2], p.shape[-1]
            if (self.nx, self.ny) != (nx, ny):
                self.create_grids(img_size, (nx, ny), p.device, p.dtype)
            return p.view(bs, self.na, self.nc + 6, self.ny, self.nx).permute(0, 1, 3, 4, 2).contiguous()
        if "default" not in self.arc:
            raise NotImplementedError("only the 'default' arcs are decoded by the fused kernel")
        p = p.float().contiguous()
        bs, ny, nx = p.shape[0], p.shape[-2], p.shape[-1]
        rows = self.na * ny * nx
        io = torch.empty((bs, rows, self.nc + 6), dtype=torch.float32, device=p.device)
        pp = self.decode_into(p, img_size, io, rows, 0)
        return io, pp


def create_modules(module_defs, arc, hyp):
    """reference model/models.py:36-164 -- same module structure and names; smart bias init omitted (it always
    fails in the reference, models.py:132-155)."""
    hyperparams = module_defs.pop(0)
    output_filters = [int(hyperparams["channels"])]
    module_list = nn.ModuleList()
    routes = []
    yolo_index = -1
    filters = output_filters[-1]
    for i, mdef in enumerate(module_defs):
        modules = nn.Sequential()
        t = mdef["type"]
        if t == "convolutional":
            bn = int(mdef["batch_normalize"])
            filters = int(mdef["filters"])
            k = int(mdef["size"])
            pad = (k - 1) // 2 if int(mdef["pad"]) else 0
            modules.add_module("Conv2d", nn.Conv2d(output_filters[-1], filters, k, int(mdef["stride"]), pad, bias=not bn))
            if bn:
                modules.add_module("BatchNorm2d", nn.BatchNorm2d(filters, momentum=0.1))
            if mdef["activation"] == "leaky":
                modules.add_module("activation", nn.PReLU(num_parameters=1, init=0.10))
        elif t == "maxpool":
            k, s = int(mdef["size"]), int(mdef["stride"])
            mp = nn.MaxPool2d(kernel_size=k, stride=s, padding=int((k - 1) // 2))
            if k == 2 and s == 1:
                modules.add_module("ZeroPad2d", nn.ZeroPad2d((0, 1, 0, 1)))
                modules.add_module("MaxPool2d", mp)
            else:
                modules = mp
        elif t == "upsample":
            modules = nn.Upsample(scale_factor=int(mdef["stride"]), mode="nearest")
        elif t == "se":
            modules = SELayer(int(mdef["channels"]))           # reference models.py:88-90
        elif t == "route":
            layers = [int(x) for x in mdef["layers"].split(",")]
            filters = sum(output_filters[l + 1 if l > 0 else l] for l in layers)
            routes.extend([l if l > 0 else l + i for l in layers])
        elif t == "shortcut":
            filters = output_filters[int(mdef["from"])]
            layer = int(mdef["from"])
            routes.extend([i + layer if layer < 0 else layer])
        elif t == "yolo":
            yolo_index += 1
            lo, hi = [int(v) for v in mdef["mask"].split("-")]
            modules = YOLOLayer(anchors=mdef["anchors"][lo:hi + 1], nc=int(mdef["classes"]), hyp=hyp,
                                yolo_index=yolo_index, arc=arc)
        else:
            raise NotImplementedError("layer type %r is outside the hot-path scope (SURVEY.md section 2)" % t)
        module_list.append(modules)
        output_filters.append(filters)
    return module_list, routes


class _View:
    """a [B, H, W, C] activation living in a (possibly wider) padded-NHWC buffer"""

    def __init__(self, buf, ch_off, c, h, w):
        self.buf, self.ch_off, self.c, self.h, self.w = buf, ch_off, c, h, w

    @property
    def cs(self):
        return self.buf.shape[-1] if self.buf is not None else 0

    @property
    def ptr(self):          # buf None: the activation exists only in a consumer's space-to-depth buffer
        return self.buf.data_ptr() + 2 * self.ch_off if self.buf is not None else None


class Darknet(nn.Module):
    def __init__(self, cfg, hyp, arc="default", precision="bf16"):
        """precision: "bf16" = throughput mode (bf16 operands / activations, fp32 accumulate); "parity" = fp32-grade
        arithmetic on the same tensor pipe (split bf16 operands, parity_path.py) for the reference's 1e-4 tolerance."""
        super().__init__()
        if precision not in ("bf16", "parity"):
            raise ValueError("precision must be 'bf16' or 'parity'")
        self.precision = precision
        self._pplan = None
        self._pplan_key = None
        self.module_defs = parse_model_cfg(cfg)
        self.net_hyper = self.module_defs[0]
        self.module_list, self.routes = create_modules(self.module_defs, arc, hyp)
        self.yolo_layers = [i for i, d in enumerate(self.module_defs) if d["type"] == "yolo"]
        self.hyp = hyp
        self.arc = arc
        self.version = np.array([0, 2, 5], dtype=np.int32)
        self.seen = np.array([0], dtype=np.int64)
        self._plan = None
        self._plan_key = None
        self._ver_tensors = None
        self._tplan = None
        self._tplan_key = None
        # opt-in: replay the eval forward as ONE CUDA graph (78 launches -> 1).  The returned tensors are then static
        # buffers that the next forward overwrites (standard CUDA-graph semantics), hence not the default.
        self.use_cuda_graph = False
        self._graph = None
        self.replayed_kernels = 0

    # ------------------------------------------------------------------------------------------------------
    def fuse(self):
        """reference models.py:300-313 folds BN into the convs module-by-module; here folding happens when the
        GEMM operands are packed (every eval forward uses folded weights), so this only drops the cache."""
        self._plan = None
        self._graph = None

    def train(self, mode=True):
        self._plan = None
        self._graph = None
        return super().train(mode)

    def load_state_dict(self, *a, **k):
        self._plan = None
        self._graph = None
        return super().load_state_dict(*a, **k)

    def _apply(self, fn, *a, **k):      # .to() / .cuda() / .half() / .float(): storage changes without a version bump
        self._plan = None
        self._graph = None
        self._tplan = None
        self._pplan = None
        self._ver_tensors = None
        return super()._apply(fn, *a, **k)

    # ------------------------------------------------------------------------------------------------------
    def _folded(self, i, device):
        """(weight [cout,cin,k,k], per-filter scale or None, bias [cout], slope or None) with eval-mode BN folded
        per utils/torch_utils.py:45-69."""
        seq = self.module_list[i]
        conv = seq.Conv2d
        w = conv.weight.detach().float()
        if hasattr(seq, "BatchNorm2d"):
            bn = seq.BatchNorm2d
            scale = bn.weight.detach().float() / torch.sqrt(bn.running_var.detach().float() + bn.eps)
            bias = bn.bias.detach().float() - bn.running_mean.detach().float() * scale
            if conv.bias is not None:
                bias = bias + conv.bias.detach().float() * scale
        else:
            scale = None
            bias = conv.bias.detach().float() if conv.bias is not None else torch.zeros(w.shape[0], device=device)
        slope = float(seq.activation.weight.detach().float().item()) if hasattr(seq, "activation") else None
        return w, scale, bias, slope

    def _build_plan(self, batch, height, width, device):
        defs = self.module_defs
        n = len(defs)
        # ---- static shape walk ----
        shape = [None] * n   # (C, H, W) of every block's output
        c, h, w = 3, height, width
        for i, d in enumerate(defs):
            t = d["type"]
            if t == "convolutional":
                s = int(d["stride"])
                c, h, w = int(d["filters"]), (h + s - 1) // s, (w + s - 1) // s
            elif t == "upsample":
                h, w = h * int(d["stride"]), w * int(d["stride"])
            elif t == "maxpool":
                if int(d["size"]) != 2 or int(d["stride"]) not in (1, 2):
                    raise NotImplementedError("only the 2x2 max pools of yolov3-tiny have a kernel")
                if int(d["stride"]) == 2:
                    h, w = h // 2, w // 2
            elif t == "route":
                ls = [int(x) for x in d["layers"].split(",")]
                ls = [l if l > 0 else i + l for l in ls]
                c = sum(shape[l][0] for l in ls)
                h, w = shape[ls[0]][1], shape[ls[0]][2]
            elif t in ("shortcut", "yolo", "se"):
                pass
            else:
                raise NotImplementedError("block type %r has no sm_100a kernel yet" % t)
            shape[i] = (c, h, w)
        # ---- concat groups: route blocks with >1 source write-share one buffer ----
        target = {}   # materialised layer index -> (buffer, channel offset)
        views = [None] * n
        for i, d in enumerate(defs):
            if d["type"] == "route":
                ls = [int(x) for x in d["layers"].split(",")]
                ls = [l if l > 0 else i + l for l in ls]
                if len(ls) > 1:
                    ctot = sum(shape[l][0] for l in ls)
                    for l in ls:
                        if shape[l][0] % 8:
                            raise NotImplementedError("concat source with channels % 8 != 0")
                    buf = L.alloc_padded(batch, shape[i][1], shape[i][2], L.round_up(ctot, 64), device)
                    off = 0
                    for l in ls:
                        target[l] = (buf, off)
                        off += shape[l][0]
                    views[i] = _View(buf, 0, ctot, shape[i][1], shape[i][2])

        def out_view(layer):
            cc, hh, ww = shape[layer]
            if layer in target:
                buf, off = target[layer]
                return _View(buf, off, cc, hh, ww)
            return _View(L.alloc_padded(batch, hh, ww, L.round_up(cc, 32), device), 0, cc, hh, ww)

        steps = []
        heads = []
        first_xs = None
        i = 0
        while i < n:
            d = defs[i]
            t = d["type"]
            if t == "convolutional":
                nxt = defs[i + 1]["type"] if i + 1 < n else None
                cin = 3 if i == 0 else shape[i - 1][0] if defs[i - 1]["type"] != "yolo" else None
                src = None if i == 0 else views[i - 1]
                cout, oh, ow = shape[i]
                k, s = int(d["size"]), int(d["stride"])
                wt, scale, bias, slope = self._folded(i, device)
                if i == 0:
                    if not (k == 3 and s == 1 and wt.shape[1] == 3 and cout in (16, 32) and slope is not None):
                        raise NotImplementedError("first layer must be 3x3/1 conv, 3 -> 16|32 channels, leaky")
                    wf = (wt * scale.view(-1, 1, 1, 1)).contiguous() if scale is not None else wt.contiguous()
                    nd = defs[1] if n > 1 else {}
                    first_s2d = (nd.get("type") == "convolutional" and int(nd["size"]) == 3 and int(nd["stride"]) == 2
                                 and height % 2 == 0 and width % 2 == 0 and 0 not in self.routes)
                    v = out_view(0) if not first_s2d else _View(None, 0, cout, height, width)
                    if first_s2d:
                        # layer 0 feeds only the stride-2 layer 1: write its output directly in that layer's
                        # space-to-depth layout (no [B, H+2, W+2, 32] tensor, no s2d pass over it)
                        first_xs = L.alloc_padded(batch, height // 2, width // 2, L.round_up(4 * cout, 64), device)
                    steps.append(("first", dict(w=wf, b=bias.contiguous(), cout=cout, slope=slope, out=v,
                                                xs=first_xs if first_s2d else None)))
                    views[0] = v
                    i += 1
                    continue
                fuse_res = nxt == "shortcut" and i not in self.routes
                fuse_up = nxt == "upsample" and i not in self.routes and int(defs[i + 1]["stride"]) == 2
                # head = the linear, BN-less conv feeding a YOLO layer (or, in YOLO-less trunk graphs, any such conv)
                is_head = nxt == "yolo" or (not self.yolo_layers and slope is None and not hasattr(self.module_list[i], "BatchNorm2d"))
                mat = i + 1 if (fuse_res or fuse_up) else i       # the layer index whose output is materialised
                if src.c != wt.shape[1]:
                    raise RuntimeError("channel mismatch at block %d" % i)
                x_ptr = src.ptr
                use_s2d = s == 2 and k == 3 and src.h % 2 == 0 and src.w % 2 == 0 and src.c % 8 == 0
                if use_s2d:
                    # 3x3/stride-2 as a 2x2-tap stride-1 conv on the space-to-depth copy of the input (1.78x instead of
                    # 4x MMA work; also removes the 32->64 channel padding of the first down-sampling layer)
                    if i == 1 and first_xs is not None:
                        xs = first_xs                      # already written by the first-layer kernel
                    else:
                        xs = L.alloc_padded(batch, src.h // 2, src.w // 2, L.round_up(4 * src.c, 64), device)
                        steps.append(("s2d", dict(x=src.ptr, xcs=src.cs, h=src.h, w=src.w, c=src.c, xs=xs, keep=src)))
                    desc = L.make_desc(batch, src.h // 2, src.w // 2, 4 * src.c, xs.shape[-1], cout, 0, 2, 1,
                                       slope is not None, slope if slope is not None else 0.0, fuse_res, 0, fuse_up, is_head)
                    wt = L.s2d_weight(wt)
                    x_ptr = xs.data_ptr()
                else:
                    desc = L.make_desc(batch, src.h, src.w, src.c, src.cs, cout, 0, k, s, slope is not None,
                                       slope if slope is not None else 0.0, fuse_res, 0, fuse_up, is_head)
                res = None
                if is_head:
                    out = torch.empty((batch, cout, oh, ow), dtype=torch.float32, device=device)
                    out_ptr = out.data_ptr()
                    heads.append((i, out))
                else:
                    v = out_view(mat)
                    desc.cout_stride = v.cs
                    out_ptr = v.ptr
                    views[mat] = v
                    views[i] = v if mat == i else None
                    if fuse_res:
                        frm = int(defs[i + 1]["from"])
                        res = views[i + 1 + frm if frm < 0 else frm]
                        if res is None or (res.c, res.h, res.w) != (cout, oh, ow):
                            raise NotImplementedError("shortcut source layout")
                        desc.res_stride = res.cs
                pw = L.pack_weights(desc, wt, scale)
                pb = L.padded_bias(desc, bias)
                steps.append(("conv", dict(desc=desc, x=x_ptr, w=pw, b=pb, y=out_ptr,
                                           r=res.ptr if res is not None else None,
                                           keep=(src, res, views[mat] if not is_head else out))))
                i += 2 if (fuse_res or fuse_up) else 1
                continue
            if t == "se":
                # squeeze-and-excitation rescales the previous block's output IN PLACE; legal when nothing else reads the
                # unscaled tensor (the reference's graphs: [se] directly follows the down-sampling conv)
                src = views[i - 1]
                if src is None or (i - 1) in self.routes or src.c != int(d["channels"]):
                    raise NotImplementedError("[se] block whose input is routed elsewhere / channel mismatch")
                fc = self.module_list[i].fc
                steps.append(("se", dict(x=src.ptr, xcs=src.cs, h=src.h, w=src.w, c=src.c,
                                         w1=fc[0].weight.detach().float().contiguous(),
                                         w2=fc[2].weight.detach().float().contiguous(), cr=fc[0].weight.shape[0],
                                         sums=torch.zeros(batch * src.c, dtype=torch.float32, device=device),
                                         scale=torch.zeros(batch * src.c, dtype=torch.float32, device=device), keep=src)))
                views[i] = src
                i += 1
                continue
            if t == "maxpool":
                src = views[i - 1]
                v = out_view(i)
                steps.append(("maxpool", dict(x=src.ptr, xcs=src.cs, h=src.h, w=src.w, c=src.c, stride=int(d["stride"]),
                                              y=v.ptr, ycs=v.cs, keep=(src, v))))
                views[i] = v
            elif t == "route":
                ls = [int(x) for x in d["layers"].split(",")]
                ls = [l if l > 0 else i + l for l in ls]
                if len(ls) == 1:
                    views[i] = views[ls[0]]
                # multi-source: views[i] was set above and is filled by its producers
            elif t == "yolo":
                views[i] = None
            elif t in ("shortcut", "upsample"):
                raise NotImplementedError("%s at block %d does not follow a convolution it can be fused into" % (t, i))
            i += 1
        rows = [self.module_list[j].na * shape[j][1] * shape[j][2] for j in self.yolo_layers]
        return dict(steps=steps, heads=heads, rows=rows, img=(height, width))

    # ------------------------------------------------------------------------------------------------------
    def forward(self, x, var=None):
        if not x.is_cuda:
            raise RuntimeError("Darknet.forward needs a CUDA tensor (sm_100a kernels, no CPU fallback)")
        x = x.float().contiguous()
        b, _, h, w = x.shape
        key = (b, h, w, x.device)
        if self.precision == "parity":
            from .parity_path import DarknetParityFn, ParityPlan
            with torch.cuda.device(x.device):
                pkey = key + (self.training,)
                if self._pplan is None or self._pplan_key != pkey:
                    self._pplan = ParityPlan(self, b, h, w, x.device, self.training)
                    self._pplan_key = pkey
                if self.training:
                    names = [n for n, _ in self.named_parameters()]
                    return list(DarknetParityFn.apply(self._pplan, x, names, *list(self.parameters())))
                with torch.no_grad():
                    heads = self._pplan.forward(x)
                    if not self.yolo_layers:
                        return [hd.clone() for hd in heads]
                    return self._decode_heads(heads, b, h, w, x.device)
        if self.training:
            # reference models.py:192-194, 290-292: list of raw [B, na, ny, nx, nc+6] tensors; differentiable
            from .train_path import DarknetTrainFn, TrainPlan
            with torch.cuda.device(x.device):
                if self._tplan is None or self._tplan_key != key:
                    self._tplan = TrainPlan(self, b, h, w, x.device)
                    self._tplan_key = key
                raw = DarknetTrainFn.apply(self._tplan, x, *list(self.parameters()))     # fp32 NCHW head buffers
                outs = []
                for t, yi in zip(raw, self.yolo_layers):
                    layer = self.module_list[yi]
                    # model/models.py:190-192 without the .contiguous(): same values, the loss reads the view in place
                    outs.append(t.view(b, layer.na, layer.nc + 6, t.shape[2], t.shape[3]).permute(0, 1, 3, 4, 2))
                return outs
        with torch.cuda.device(x.device):
            # the plan holds PACKED (BN-folded, bf16) copies of the weights: rebuild it when any parameter or buffer was
            # modified in place since (optimizer step, manual edits); train()/eval()/load_state_dict drop it explicitly
            if self._ver_tensors is None:      # walking the module tree costs ~1 ms; the tensor objects are stable
                self._ver_tensors = list(self.parameters()) + list(self.buffers())
            key = key + (sum(t._version for t in self._ver_tensors),)
            if self._plan is None or self._plan_key != key:
                self._plan = self._build_plan(b, h, w, x.device)
                self._plan_key = key
                self._graph = None
            if self.use_cuda_graph:
                if self._graph is None:
                    self._x_static = x.clone()
                    self._run_eval(self._x_static, b, h, w)          # warm-up (attribute calls, lazy inits) outside capture
                    torch.cuda.synchronize()
                    g = torch.cuda.CUDAGraph()
                    c0 = _lib.lib.ryolo_launch_count()
                    with torch.cuda.graph(g):
                        self._graph_out = self._run_eval(self._x_static, b, h, w)
                    self._graph = g
                    self._graph_kernels = int(_lib.lib.ryolo_launch_count() - c0)
                self._x_static.copy_(x)
                self._graph.replay()
                self.replayed_kernels += self._graph_kernels     # this library's kernels executed by the replay
                return self._graph_out
            return self._run_eval(x, b, h, w)

    def _run_eval(self, x, b, h, w):
        plan = self._plan
        lib = _lib.lib
        stream = _lib.stream_ptr(x.device)
        for kind, a in plan["steps"]:
            if kind == "maxpool":
                st = lib.ryolo_maxpool2x2(ctypes.c_void_p(a["x"]), a["xcs"], b, a["h"], a["w"], a["c"], a["stride"],
                                          ctypes.c_void_p(a["y"]), a["ycs"], stream)
                _lib.check(st, "ryolo_maxpool2x2")
            elif kind == "se":
                st = lib.ryolo_se_block(ctypes.c_void_p(a["x"]), a["xcs"], b, a["h"], a["w"], a["c"], _lib.ptr(a["w1"]),
                                        _lib.ptr(a["w2"]), a["cr"], _lib.ptr(a["sums"]), _lib.ptr(a["scale"]), stream)
                _lib.check(st, "ryolo_se_block")
            elif kind == "s2d":
                st = lib.ryolo_space_to_depth(ctypes.c_void_p(a["x"]), a["xcs"], b, a["h"], a["w"], a["c"],
                                              _lib.ptr(a["xs"]), a["xs"].shape[-1], stream)
                _lib.check(st, "ryolo_space_to_depth")
            elif kind == "first":
                if a["xs"] is not None:
                    st = lib.ryolo_conv_first_s2d_fwd(_lib.ptr(x), b, h, w, _lib.ptr(a["w"]), _lib.ptr(a["b"]),
                                                      a["cout"], a["slope"], _lib.ptr(a["xs"]), a["xs"].shape[-1], stream)
                else:
                    v = a["out"]
                    st = lib.ryolo_conv_first_fwd(_lib.ptr(x), b, h, w, _lib.ptr(a["w"]), _lib.ptr(a["b"]), a["cout"],
                                                  a["slope"], ctypes.c_void_p(v.ptr), v.cs, stream)
                _lib.check(st, "ryolo_conv_first_fwd")
            else:
                st = lib.ryolo_conv_bn_act_fwd(ctypes.byref(a["desc"]), ctypes.c_void_p(a["x"]), _lib.ptr(a["w"]),
                                               _lib.ptr(a["b"]), ctypes.c_void_p(a["r"]) if a["r"] else None,
                                               ctypes.c_void_p(a["y"]), None, 0, stream)
                _lib.check(st, "ryolo_conv_bn_act_fwd")
        if not self.yolo_layers:      # trunk graph without YOLO layers: the raw head maps [B, filters, ny, nx]
            return [head for _, head in plan["heads"]]
        return self._decode_heads([head for _, head in plan["heads"]], b, h, w, x.device)

    def _decode_heads(self, heads, b, h, w, device):
        """fp32 NCHW head maps -> (io [B, sum(na*ny*nx), nc+6], (p0, p1, p2)) like models.py:294-298"""
        nc = self.module_list[self.yolo_layers[0]].nc
        rows = [self.module_list[yi].na * hd.shape[2] * hd.shape[3] for yi, hd in zip(self.yolo_layers, heads)]
        total = sum(rows)
        io = torch.empty((b, total, nc + 6), dtype=torch.float32, device=device)
        ps = []
        off = 0
        for head, yi, r in zip(heads, self.yolo_layers, rows):
            ps.append(self.module_list[yi].decode_into(head, (h, w), io, total, off))
            off += r
        return io, tuple(ps)
