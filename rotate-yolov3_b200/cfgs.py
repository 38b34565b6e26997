"""Generators for the Darknet block graphs on the hot path (no reference file is copied; the graphs are the
public Darknet-53 / YOLOv3 and YOLOv3-tiny topologies).

yolov3_cfg()      same layer graph as the reference's cfg/yolov3.cfg (107 blocks): 75 convs, 23 shortcuts, 4 routes,
                  2 upsamples, 3 rotated YOLO heads with na = areas*ratios*angles/3 anchors per scale.
yolov3_tiny_trunk_cfg()  the 13-conv / 6-maxpool trunk of cfg/yolov3-tiny.cfg (BASELINE configs[0])."""

DEFAULT_ANCHORS = ("792, 2061, 3870, 6353, 9623, 15803 / 4.18, 6.48, 8.71  / "
                   "-75, -60, -45, -30, -15 ,0,15, 30,45, 60,75, 90")


def _conv(filters, size, stride=1, bn=1, act="leaky"):
    s = "[convolutional]\n"
    if bn:
        s += "batch_normalize=1\n"
    return s + "filters=%d\nsize=%d\nstride=%d\npad=1\nactivation=%s\n\n" % (filters, size, stride, act)


def _res(n, c):
    s = ""
    for _ in range(n):
        s += _conv(c // 2, 1) + _conv(c, 3) + "[shortcut]\nfrom=-3\nactivation=linear\n\n"
    return s


def yolov3_cfg(width=608, height=608, classes=1, anchors=DEFAULT_ANCHORS, n_anchors=216):
    na = n_anchors // 3
    head_filters = na * (classes + 6)

    def yolo(lo, hi):
        return ("[yolo]\nmask = %d-%d\nanchors = %s\nclasses=%d\nnum=%d\njitter=.3\nignore_thresh = .7\n"
                "truth_thresh = 1\nrandom=1\n\n" % (lo, hi, anchors, classes, n_anchors))

    s = "[net]\nbatch=64\nsubdivisions=16\nwidth=%d\nheight=%d\nchannels=3\n\n" % (width, height)
    s += _conv(32, 3) + _conv(64, 3, 2) + _res(1, 64)
    s += _conv(128, 3, 2) + _res(2, 128)
    s += _conv(256, 3, 2) + _res(8, 256)
    s += _conv(512, 3, 2) + _res(8, 512)
    s += _conv(1024, 3, 2) + _res(4, 1024)
    for _ in range(3):
        s += _conv(512, 1) + _conv(1024, 3)
    s += _conv(head_filters, 1, bn=0, act="linear") + yolo(2 * na, 3 * na - 1)
    s += "[route]\nlayers = -4\n\n" + _conv(256, 1) + "[upsample]\nstride=2\n\n[route]\nlayers = -1, 61\n\n"
    for _ in range(3):
        s += _conv(256, 1) + _conv(512, 3)
    s += _conv(head_filters, 1, bn=0, act="linear") + yolo(na, 2 * na - 1)
    s += "[route]\nlayers = -4\n\n" + _conv(128, 1) + "[upsample]\nstride=2\n\n[route]\nlayers = -1, 36\n\n"
    for _ in range(3):
        s += _conv(128, 1) + _conv(256, 3)
    s += _conv(head_filters, 1, bn=0, act="linear") + yolo(0, na - 1)
    return s


def yolov3_tiny_trunk_cfg(width=416, height=416):
    s = "[net]\nbatch=1\nwidth=%d\nheight=%d\nchannels=3\n\n" % (width, height)
    for i, c in enumerate((16, 32, 64, 128, 256, 512)):
        s += _conv(c, 3) + "[maxpool]\nsize=2\nstride=%d\n\n" % (2 if i < 5 else 1)
    s += _conv(1024, 3) + _conv(256, 1) + _conv(512, 3) + _conv(255, 1, bn=0, act="linear")
    s += "[route]\nlayers = -4\n\n" + _conv(128, 1) + "[upsample]\nstride=2\n\n[route]\nlayers = -1, 8\n\n"
    s += _conv(256, 3) + _conv(255, 1, bn=0, act="linear")
    return s
