#!/usr/bin/env python
"""Golden fixtures at the BASELINE shape: the REFERENCE Darknet (model/models.py) built from the cfg/yolov3.cfg graph
with its 216-anchor line (504-channel heads) at 608 x 608, run here on CPU in fp32.

    darknet608_golden.npz   eval forward, batch 1 (randomised BN statistics / slopes): sub-sampled head tensors and
                            decoded rows; training-mode forward + backward, batch 2: sub-sampled head tensors, gradient
                            samples + norms of every parameter (more samples for the three head convs).

Inputs, weights, the loss cotangents and the sample positions are all regenerated from seeds by the test
(tests/test_parity_gpu.py) -- the fixture stores only reference OUTPUTS plus a few input samples as a cross-check.
Not run by the test-suite (the GPU box has no /root/reference).  Re-run: python tests/golden/make_golden_608.py"""
import os
import sys
import tempfile

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.join(REPO, "tests"))
sys.path.insert(0, REPO)

N_SAMPLE = 100000


def sample_idx(numel, n, seed):
    g = torch.Generator().manual_seed(seed)
    return torch.randint(0, numel, (min(n, numel),), generator=g)


def main():
    import make_golden as mg
    import helpers
    import importlib.util
    spec = importlib.util.spec_from_file_location("cfgs", os.path.join(REPO, "rotate-yolov3_b200", "cfgs.py"))
    cfgs = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(cfgs)
    ru, rnms, rmodels, rn_stub = mg.import_reference()
    torch.set_num_threads(max(1, len(os.sched_getaffinity(0))))
    # 608 x 608, 216 anchors, 1 class.  The reference parser needs the "ara" token that cfg/yolov3.cfg itself lacks
    text = cfgs.yolov3_cfg(anchors="ara " + cfgs.DEFAULT_ANCHORS)
    with tempfile.NamedTemporaryFile("w", suffix=".cfg", delete=False) as f:
        f.write(text)
        cfg_path = f.name
    sav = {}
    # ---- eval, batch 1 ----
    model = rmodels.Darknet(cfg_path, {"context_factor": 1.0}, arc="default")
    helpers.init_darknet_weights(model, seed=608)
    model.eval()
    x = torch.rand(1, 3, 608, 608, generator=torch.Generator().manual_seed(6080))
    with torch.no_grad():
        io, ps = model(x)
    sav["x_probe"] = x.reshape(-1)[:64].numpy()
    sav["n_params"] = sum(p.numel() for p in model.parameters())
    for k, p in enumerate(ps):
        flat = p.reshape(-1)
        idx = sample_idx(flat.numel(), N_SAMPLE, 100 + k)
        sav["eval_p%d" % k] = flat[idx].numpy()
        sav["eval_p%d_absmax" % k] = float(flat.abs().max())
        sav["eval_p%d_rms" % k] = float(flat.pow(2).mean().sqrt())
    rows = sample_idx(io.shape[1], 20000, 110)
    sav["eval_io_rows"] = io[0, rows].numpy()
    print("eval:", tuple(io.shape), [tuple(p.shape) for p in ps], [sav["eval_p%d_absmax" % k] for k in range(3)])

    # ---- training mode, batch 2: forward + backward through a seeded linear functional of the heads ----
    model = rmodels.Darknet(cfg_path, {"context_factor": 1.0}, arc="default")
    helpers.init_darknet_weights(model, seed=609)
    model.train()
    x = torch.rand(2, 3, 608, 608, generator=torch.Generator().manual_seed(6090))
    ps = model(x)
    g = torch.Generator().manual_seed(6091)
    gs = [torch.randn(p.shape, generator=g) for p in ps]
    loss = sum((p * gg).sum() for p, gg in zip(ps, gs)) / 100.0
    loss.backward()
    sav["train_loss"] = float(loss)
    for k, p in enumerate(ps):
        flat = p.detach().reshape(-1)
        idx = sample_idx(flat.numel(), N_SAMPLE, 200 + k)
        sav["train_p%d" % k] = flat[idx].numpy()
        sav["train_p%d_absmax" % k] = float(flat.abs().max())
    names, norms, absmax, samples = [], [], [], []
    for j, (name, prm) in enumerate(model.named_parameters()):
        gflat = prm.grad.reshape(-1)
        n = 8192 if name.split(".")[1] in ("81", "93", "105") else 128
        idx = sample_idx(gflat.numel(), n, 1000 + j)
        names.append(name)
        norms.append(float(gflat.norm()))
        absmax.append(float(gflat.abs().max()))
        samples.append(gflat[idx].numpy())
    sav["grad_names"] = np.array(names)
    sav["grad_norms"] = np.array(norms)
    sav["grad_absmax"] = np.array(absmax)
    sav["grad_samples"] = np.array(samples, dtype=object)
    bn0 = model.module_list[0].BatchNorm2d
    sav["rm0"], sav["rv0"] = bn0.running_mean.numpy(), bn0.running_var.numpy()
    sav["rm104"] = model.module_list[104].BatchNorm2d.running_mean.numpy()
    np.savez_compressed(os.path.join(HERE, "darknet608_golden.npz"), **sav)
    print("train: loss %.5f, %d params, grad norm range %.3e .. %.3e" % (float(loss), len(names), min(norms), max(norms)))
    print("wrote", os.path.join(HERE, "darknet608_golden.npz"), os.path.getsize(os.path.join(HERE, "darknet608_golden.npz")))


if __name__ == "__main__":
    main()
