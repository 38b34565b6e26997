"""Darknet .weights I/O against a file written by the reference's own save_weights (tests/golden/make_golden.py 10)."""
import filecmp
import os

import numpy as np
import torch

from helpers import GOLDEN, init_darknet_weights


def test_load_reference_file_and_write_identical_bytes(tmp_path):
    import rotate_yolov3_b200 as pkg
    from rotate_yolov3_b200 import weights_io
    cfg = open(os.path.join(GOLDEN, "micro.cfg")).read()
    want = pkg.Darknet(cfg, {"context_factor": 1.0})
    init_darknet_weights(want, seed=9)                     # what the reference model held when it wrote the file
    m = pkg.Darknet(cfg, {"context_factor": 1.0})
    weights_io.load_darknet_weights(m, os.path.join(GOLDEN, "micro_reference.weights"))
    assert int(m.seen[0]) == 1234 and list(m.version) == [0, 2, 5]
    for (n, a), (_, b) in zip(m.state_dict().items(), want.state_dict().items()):
        if "activation" in n or "num_batches" in n:
            continue                                       # PReLU slopes are not part of the format
        assert torch.equal(a, b), n
    out = tmp_path / "roundtrip.weights"
    weights_io.save_weights(m, str(out))
    assert filecmp.cmp(str(out), os.path.join(GOLDEN, "micro_reference.weights"), shallow=False)


def test_reference_checkpoint_loads_and_round_trips(tmp_path):
    """a .pt written by the reference's train.py code path (dict layout of train.py:349-355, reference Darknet state_dict,
    SGD momentum buffers) loads into this Darknet + optimizer; save_checkpoint writes the same layout back; convert()
    goes .pt -> .weights -> .pt"""
    import rotate_yolov3_b200 as pkg
    from rotate_yolov3_b200 import weights_io
    cfg = open(os.path.join(GOLDEN, "micro.cfg")).read()
    ref = torch.load(os.path.join(GOLDEN, "micro_reference.pt"), map_location="cpu", weights_only=False)
    m = pkg.Darknet(cfg, {"context_factor": 1.0})
    assert list(m.state_dict().keys()) == list(ref["model"].keys())          # same names, same order
    pg0, pg1 = [], []
    for k, v in dict(m.named_parameters()).items():
        (pg1 if "Conv2d.weight" in k else pg0).append(v)
    opt = torch.optim.SGD(pg0, lr=1e-3, momentum=0.9, nesterov=True)
    opt.add_param_group({"params": pg1, "weight_decay": 5e-4})
    start_epoch, best, results = weights_io.load_checkpoint(m, os.path.join(GOLDEN, "micro_reference.pt"), opt, resume=True)
    assert start_epoch == 8 and abs(best - 0.4242) < 1e-12 and results == "epoch 7 results line\n"
    for k, v in ref["model"].items():
        assert torch.equal(m.state_dict()[k], v), k
    assert len(opt.state_dict()["state"]) == len(ref["optimizer"]["state"]) > 0
    for k, st in ref["optimizer"]["state"].items():
        assert torch.equal(opt.state_dict()["state"][k]["momentum_buffer"], st["momentum_buffer"])
    # write it back: same layout, same tensors
    out = tmp_path / "last.pt"
    weights_io.save_checkpoint(str(out), m, opt, epoch=7, best_fitness=best, training_results=results)
    again = torch.load(str(out), map_location="cpu", weights_only=False)
    assert set(again) == set(ref) and again["epoch"] == 7
    assert all(torch.equal(again["model"][k], ref["model"][k]) for k in ref["model"])
    assert weights_io.save_checkpoint(str(out), m, opt, final_epoch=True)["optimizer"] is None     # train.py:355
    # transfer: a checkpoint tensor with another element count is skipped, not an error (train.py:99-100)
    other = dict(ref)
    other["model"] = dict(ref["model"])
    other["model"]["module_list.2.Conv2d.weight"] = torch.zeros(21, 16, 1, 1)
    torch.save(other, str(tmp_path / "other.pt"))
    m2 = pkg.Darknet(cfg, {"context_factor": 1.0})
    before = m2.state_dict()["module_list.2.Conv2d.weight"].clone()
    weights_io.load_checkpoint(m2, str(tmp_path / "other.pt"))
    assert torch.equal(m2.state_dict()["module_list.2.Conv2d.weight"], before)
    assert torch.equal(m2.state_dict()["module_list.0.Conv2d.weight"], ref["model"]["module_list.0.Conv2d.weight"])
    # convert: .pt -> .weights (byte-identical to what the reference's save_weights wrote for the same parameters, up to the
    # header's `seen`) -> .pt
    wpath = weights_io.convert(cfg, os.path.join(GOLDEN, "micro_reference.pt"), out=str(tmp_path / "c.weights"))
    m3 = pkg.Darknet(cfg, {"context_factor": 1.0})
    weights_io.load_darknet_weights(m3, wpath)
    for k, v in ref["model"].items():
        if "activation" in k or "num_batches" in k:
            continue
        assert torch.equal(m3.state_dict()[k], v), k
    ppath = weights_io.convert(cfg, wpath, out=str(tmp_path / "c.pt"))
    back = torch.load(ppath, map_location="cpu", weights_only=False)
    assert back["epoch"] == -1 and back["optimizer"] is None and set(back["model"]) == set(ref["model"])
