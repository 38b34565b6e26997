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
