"""BASELINE configs[0]: the yolov3-tiny conv trunk (13 convs, 6 max-pools incl. the size-2/stride-1 zero-pad case, one
upsample + concat) on a 1 x 3 x 416 x 416 random tensor.  The reference cannot build cfg/yolov3-tiny.cfg (stock-darknet
anchor grammar, 255-channel heads), so the oracle is the plain PyTorch evaluation of the same module list
(reference semantics of create_modules / Darknet.forward) with shared weights, on the CPU."""
import pytest
import torch

from helpers import init_darknet_weights

pytestmark = pytest.mark.gpu


def test_tiny_trunk_vs_torch_modules():
    import bench
    import rotate_yolov3_b200 as pkg
    from rotate_yolov3_b200 import cfgs
    m = pkg.Darknet(cfgs.yolov3_tiny_trunk_cfg(416, 416), {"context_factor": 1.0}, arc="default")
    init_darknet_weights(m, seed=5)
    torch.manual_seed(0)
    x = torch.rand(1, 3, 416, 416)
    m.eval()
    with torch.no_grad():
        want = bench.torch_port_forward(m, x, False)
    assert [tuple(t.shape) for t in want] == [(1, 255, 13, 13), (1, 255, 26, 26)]
    m = m.cuda()
    with torch.no_grad():
        got = m(x.cuda())
    assert [tuple(t.shape) for t in got] == [(1, 255, 13, 13), (1, 255, 26, 26)]
    for g, w in zip(got, want):
        err = (g.cpu() - w).abs()
        scale = float(w.abs().max())
        # bf16 operands through 13 conv layers vs fp32: 2e-2 of the output scale (max), 4e-3 rms
        assert float(err.max()) <= 2e-2 * scale and float(err.pow(2).mean().sqrt()) <= 4e-3 * scale, (float(err.max()), scale)


def test_maxpool_kernels_vs_torch():
    import rotate_yolov3_b200 as pkg
    from rotate_yolov3_b200 import layout as L
    import torch.nn as nn
    dev = torch.device("cuda")
    x = torch.randn(2, 32, 12, 10).to(dev).to(torch.bfloat16).float()      # negative values: the zero pad matters
    xb = L.to_padded_nhwc(x, 64)
    for stride, ref in ((2, nn.MaxPool2d(2, 2)), (1, nn.Sequential(nn.ZeroPad2d((0, 1, 0, 1)), nn.MaxPool2d(2, 1)))):
        want = ref(x)
        oh, ow = want.shape[2], want.shape[3]
        y = L.alloc_padded(2, oh, ow, 64, dev)
        st = pkg._lib.lib.ryolo_maxpool2x2(pkg._lib.ptr(xb), 64, 2, 12, 10, 32, stride, pkg._lib.ptr(y), 64,
                                           pkg._lib.stream_ptr(dev))
        assert st == 0
        assert torch.equal(L.from_padded_nhwc(y, 32), want)
        assert float(y[..., 32:].abs().max()) == 0 and float(y[:, 0].abs().max()) == 0
