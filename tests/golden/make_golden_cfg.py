#!/usr/bin/env python
"""cfg_parse_golden.json: what the REFERENCE parsers (utils/parse_config.py:37-59 parse_model_cfg, :6-31 cfg2anchors,
utils/utils.py:33-47 hyp_parse) return for the reference's own files -- ara-grammar cfg, k-means-txt cfg, the
hyper-parameter files.  tests/test_parse_config.py feeds the same files (read from /root/reference when it exists, i.e.
in the build container) to rotate-yolov3_b200/parse_config.py and compares.  Only parser OUTPUT is stored, no cfg text."""
import contextlib
import io
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)

CFGS = ["cfg/HRSC+/yolov3_512_ma.cfg", "cfg/ICDAR/yolov3_608.cfg", "cfg/HRSC+/yolov3_512.cfg"]
HYPS = ["cfg/hyp_template.py", "cfg/ICDAR/hyp.py", "cfg/HRSC+/hyp.py"]


def jsonable(d):
    return {k: (np.asarray(v).tolist() if isinstance(v, np.ndarray) else v) for k, v in d.items()}


def main():
    import make_golden as mg
    ru, rnms, rmodels, rn_stub = mg.import_reference()     # chdir(/root/reference): txt anchors use relative paths
    from utils.parse_config import parse_model_cfg
    out = {"cfgs": {}, "hyps": {}}
    for c in CFGS:
        out["cfgs"][c] = [jsonable(b) for b in parse_model_cfg(c)]
    for h in HYPS:
        with contextlib.redirect_stdout(io.StringIO()):     # the reference prints the dict
            out["hyps"][h] = ru.hyp_parse(h)
    with open(os.path.join(HERE, "cfg_parse_golden.json"), "w") as f:
        json.dump(out, f)
    print({k: len(v) for k, v in out["cfgs"].items()}, {k: len(v) for k, v in out["hyps"].items()})

    # ---- a checkpoint exactly as the reference writes it (train.py:345-363): reference Darknet on tests/golden/micro.cfg,
    # one SGD step so that the optimizer state holds momentum buffers, dict layout of train.py:349-355 ----
    import torch
    sys.path.insert(0, os.path.join(os.path.dirname(HERE)))
    import helpers
    model = rmodels.Darknet(os.path.join(HERE, "micro.cfg"), {"context_factor": 1.0}, arc="default")
    helpers.init_darknet_weights(model, seed=9)
    pg0, pg1 = [], []
    for k, v in dict(model.named_parameters()).items():       # train.py:70-76
        (pg1 if "Conv2d.weight" in k else pg0).append(v)
    optimizer = torch.optim.SGD(pg0, lr=1e-3, momentum=0.9, nesterov=True)
    optimizer.add_param_group({"params": pg1, "weight_decay": 5e-4})
    model.train()
    x = torch.rand(2, 3, 32, 32, generator=torch.Generator().manual_seed(1))
    sum(p.pow(2).mean() for p in model(x)).backward()
    optimizer.step()
    chkpt = {"epoch": 7, "best_fitness": 0.4242, "training_results": "epoch 7 results line\n", "model": model.state_dict(),
             "optimizer": optimizer.state_dict()}
    torch.save(chkpt, os.path.join(HERE, "micro_reference.pt"))
    model.eval()
    with torch.no_grad():
        io_out, _ = model(x)
    np.savez_compressed(os.path.join(HERE, "micro_ckpt_golden.npz"), x=x.numpy(), io=io_out.numpy())
    print("checkpoint golden:", os.path.getsize(os.path.join(HERE, "micro_reference.pt")), "bytes", tuple(io_out.shape))


if __name__ == "__main__":
    main()
