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
