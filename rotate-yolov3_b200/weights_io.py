"""Darknet binary ``.weights`` reader / writer with the reference's layout (model/model_utils.py:38-118): header =
int32[3] version + int64[1] images seen; then per convolutional block, in graph order, [BN bias, BN weight, running
mean, running var | conv bias] followed by the conv weight, all float32.  PReLU slopes are not part of the format (the
reference loses them too).  Pure host I/O: SURVEY.md 8f item 3."""
from pathlib import Path

import numpy as np
import torch


def load_darknet_weights(model, weights, cutoff=-1):
    name = Path(weights).name
    if name == "darknet53.conv.74":
        cutoff = 75
    elif name == "yolov3-tiny.conv.15":
        cutoff = 15
    with open(weights, "rb") as f:
        model.version = np.fromfile(f, dtype=np.int32, count=3)
        model.seen = np.fromfile(f, dtype=np.int64, count=1)
        w = np.fromfile(f, dtype=np.float32)
    ptr = 0

    def take(t):
        nonlocal ptr
        n = t.numel()
        t.data.copy_(torch.from_numpy(w[ptr:ptr + n]).view_as(t))
        ptr += n
    for mdef, module in zip(model.module_defs[:cutoff], model.module_list[:cutoff]):
        if mdef["type"] != "convolutional":
            continue
        conv = module.Conv2d
        if int(mdef["batch_normalize"]):
            bn = module.BatchNorm2d
            for t in (bn.bias, bn.weight, bn.running_mean, bn.running_var):
                take(t)
        else:
            take(conv.bias)
        take(conv.weight)
    if hasattr(model, "_plan"):
        model._plan = None
        model._graph = None
    return cutoff


def save_weights(model, path="model.weights", cutoff=-1):
    with open(path, "wb") as f:
        model.version.tofile(f)
        model.seen.tofile(f)
        for mdef, module in zip(model.module_defs[:cutoff], model.module_list[:cutoff]):
            if mdef["type"] != "convolutional":
                continue
            conv = module.Conv2d
            if int(mdef["batch_normalize"]):
                bn = module.BatchNorm2d
                for t in (bn.bias, bn.weight, bn.running_mean, bn.running_var):
                    t.data.cpu().numpy().tofile(f)
            else:
                conv.bias.data.cpu().numpy().tofile(f)
            conv.weight.data.cpu().numpy().tofile(f)


# ---------------------------------------------------------------------------------------------------------------
# PyTorch checkpoints with the reference's dict layout (train.py:345-363 writes, train.py:89-115 reads)
# ---------------------------------------------------------------------------------------------------------------
def _unwrap(model):
    return model.module if hasattr(model, "module") and hasattr(model.module, "module_defs") else model


def save_checkpoint(path, model, optimizer=None, epoch=-1, best_fitness=None, training_results=None, final_epoch=False):
    """{'epoch', 'best_fitness', 'training_results', 'model': state_dict (of .module under DDP), 'optimizer': state_dict or
    None on the final epoch} -- train.py:349-355.  Parameter names are the reference's (module_list.{i}.Conv2d.weight ...),
    so the reference loads these files and vice versa."""
    chkpt = {"epoch": epoch, "best_fitness": best_fitness, "training_results": training_results,
             "model": {k: v.detach().cpu() for k, v in _unwrap(model).state_dict().items()},
             "optimizer": None if (final_epoch or optimizer is None) else optimizer.state_dict()}
    torch.save(chkpt, path)
    return chkpt


def load_checkpoint(model, path, optimizer=None, resume=False, map_location="cpu"):
    """train.py:93-115: tensors whose element count does not match the model are dropped (transfer learning across anchor
    / class counts), the rest is loaded with strict=False; the optimizer state and best fitness are restored when present.
    Returns (start_epoch, best_fitness, training_results)."""
    chkpt = torch.load(path, map_location=map_location, weights_only=False)
    net = _unwrap(model)
    own = net.state_dict()
    state = {k: v for k, v in chkpt["model"].items() if k in own and own[k].numel() == v.numel()}
    net.load_state_dict(state, strict=False)
    best_fitness = 0.0
    if optimizer is not None and chkpt.get("optimizer") is not None:
        optimizer.load_state_dict(chkpt["optimizer"])
        best_fitness = chkpt["best_fitness"]
    start_epoch = chkpt["epoch"] + 1 if resume else 0
    return start_epoch, best_fitness, chkpt.get("training_results")


def convert(cfg, weights, hyp=None, out=None):
    """model/model_utils.py:121-147: '*.pt' -> Darknet '*.weights' and back, by extension.  (The reference's own convert()
    calls Darknet(cfg) without the hyp argument its constructor requires and cannot run; same behaviour otherwise.)
    Returns the path written."""
    from .models import Darknet
    model = Darknet(cfg, hyp if hyp is not None else {"context_factor": 1.0})
    if weights.endswith(".pt"):
        model.load_state_dict(torch.load(weights, map_location="cpu", weights_only=False)["model"])
        out = out or "converted.weights"
        save_weights(model, path=out, cutoff=-1)
    elif weights.endswith(".weights"):
        load_darknet_weights(model, weights)
        out = out or "converted.pt"
        torch.save({"epoch": -1, "best_fitness": None, "training_results": None, "model": model.state_dict(),
                    "optimizer": None}, out)
    else:
        raise ValueError("extension not supported: %r" % weights)
    return out
