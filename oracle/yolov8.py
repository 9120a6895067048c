"""fp32 PyTorch restatement of the YOLOv8 detect / OBB graph (oracle; test infrastructure only).

Follows ultralytics ``cfg/models/v8/yolov8.yaml`` / ``yolov8-obb.yaml`` and ``nn/modules/{conv,block,head}.py`` as
restated in SURVEY.md Appendix A-1 (the package itself is not vendored under /root/reference; call site
/root/reference/geotrax/extract.py:153 ``model.track(frame, ...)``).  Parameter names follow the ultralytics
``state_dict`` layout (``model.<idx>.<sub>...``) so a real checkpoint's tensors map 1:1.

Pinned facts (tests/test_oracle_model.py): 11,166,560 parameters at nc=80 and 11,137,148 at nc=4 (s-scale).
"""
from __future__ import annotations

import math
from typing import Dict, List, Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F


class Conv(nn.Module):
    """Conv2d(bias=False, pad=k//2) -> BatchNorm2d(eps=1e-3) -> SiLU."""

    def __init__(self, c1, c2, k=1, s=1):
        super().__init__()
        self.conv = nn.Conv2d(c1, c2, k, s, k // 2, bias=False)
        self.bn = nn.BatchNorm2d(c2, eps=1e-3, momentum=0.03)

    def forward(self, x):
        return F.silu(self.bn(self.conv(x)))


class Bottleneck(nn.Module):
    def __init__(self, c1, c2, shortcut=True):
        super().__init__()
        self.cv1 = Conv(c1, c2, 3, 1)  # e = 1.0 inside C2f
        self.cv2 = Conv(c2, c2, 3, 1)
        self.add = shortcut and c1 == c2

    def forward(self, x):
        y = self.cv2(self.cv1(x))
        return x + y if self.add else y


class C2f(nn.Module):
    def __init__(self, c1, c2, n=1, shortcut=False):
        super().__init__()
        self.c = c2 // 2
        self.cv1 = Conv(c1, 2 * self.c, 1, 1)
        self.cv2 = Conv((2 + n) * self.c, c2, 1, 1)
        self.m = nn.ModuleList(Bottleneck(self.c, self.c, shortcut) for _ in range(n))

    def forward(self, x):
        y = list(self.cv1(x).chunk(2, 1))
        y.extend(m(y[-1]) for m in self.m)
        return self.cv2(torch.cat(y, 1))


class SPPF(nn.Module):
    def __init__(self, c1, c2, k=5):
        super().__init__()
        c_ = c1 // 2
        self.cv1 = Conv(c1, c_, 1, 1)
        self.cv2 = Conv(c_ * 4, c2, 1, 1)
        self.k = k

    def forward(self, x):
        y = [self.cv1(x)]
        for _ in range(3):
            y.append(F.max_pool2d(y[-1], self.k, 1, self.k // 2))
        return self.cv2(torch.cat(y, 1))


class DFL(nn.Module):
    def __init__(self, c1=16):
        super().__init__()
        self.conv = nn.Conv2d(c1, 1, 1, bias=False).requires_grad_(False)
        self.conv.weight.data[:] = torch.arange(c1, dtype=torch.float).view(1, c1, 1, 1)
        self.c1 = c1

    def forward(self, x):
        b, _, a = x.shape
        return self.conv(x.view(b, 4, self.c1, a).transpose(2, 1).softmax(1)).view(b, 4, a)


def make_anchors(shapes: List[Tuple[int, int]], strides, offset=0.5):
    pts, st = [], []
    for (h, w), s in zip(shapes, strides):
        sx = torch.arange(w, dtype=torch.float32) + offset
        sy = torch.arange(h, dtype=torch.float32) + offset
        yy, xx = torch.meshgrid(sy, sx, indexing="ij")
        pts.append(torch.stack((xx, yy), -1).view(-1, 2))
        st.append(torch.full((h * w, 1), float(s)))
    return torch.cat(pts).t().contiguous(), torch.cat(st).t().contiguous()  # (2,A), (1,A)


class Detect(nn.Module):
    """Legacy v8 head (cv2: box branch 4*reg_max, cv3: class branch)."""

    def __init__(self, nc, ch):
        super().__init__()
        self.nc, self.reg_max, self.nl = nc, 16, len(ch)
        self.no = nc + 4 * self.reg_max
        c2, c3 = max(16, ch[0] // 4, self.reg_max * 4), max(ch[0], min(nc, 100))
        self.cv2 = nn.ModuleList(
            nn.Sequential(Conv(x, c2, 3), Conv(c2, c2, 3), nn.Conv2d(c2, 4 * self.reg_max, 1)) for x in ch)
        self.cv3 = nn.ModuleList(
            nn.Sequential(Conv(x, c3, 3), Conv(c3, c3, 3), nn.Conv2d(c3, nc, 1)) for x in ch)
        self.dfl = DFL(self.reg_max)
        self.stride = (8.0, 16.0, 32.0)

    def raw(self, feats):
        """(B, 64+nc, A) concatenation of the per-level head outputs, levels P3,P4,P5."""
        outs = [torch.cat((self.cv2[i](f), self.cv3[i](f)), 1) for i, f in enumerate(feats)]
        shapes = [tuple(o.shape[2:]) for o in outs]
        b = outs[0].shape[0]
        return torch.cat([o.view(b, self.no, -1) for o in outs], 2), shapes

    def decode(self, raw, shapes):
        anchors, strides = make_anchors(shapes, self.stride)
        box, cls = raw.split((4 * self.reg_max, self.nc), 1)
        d = self.dfl(box)
        lt, rb = d.chunk(2, 1)
        x1y1, x2y2 = anchors.unsqueeze(0) - lt, anchors.unsqueeze(0) + rb
        dbox = torch.cat(((x1y1 + x2y2) / 2, x2y2 - x1y1), 1) * strides
        return torch.cat((dbox, cls.sigmoid()), 1)

    def forward(self, feats):
        raw, shapes = self.raw(feats)
        return self.decode(raw, shapes), raw


class OBB(Detect):
    def __init__(self, nc, ch, ne=1):
        super().__init__(nc, ch)
        self.ne = ne
        c4 = max(ch[0] // 4, ne)
        self.cv4 = nn.ModuleList(
            nn.Sequential(Conv(x, c4, 3), Conv(c4, c4, 3), nn.Conv2d(c4, ne, 1)) for x in ch)

    def forward(self, feats):
        b = feats[0].shape[0]
        angle_raw = torch.cat([self.cv4[i](f).view(b, self.ne, -1) for i, f in enumerate(feats)], 2)
        angle = (angle_raw.sigmoid() - 0.25) * math.pi
        raw, shapes = self.raw(feats)
        anchors, strides = make_anchors(shapes, self.stride)
        box, cls = raw.split((4 * self.reg_max, self.nc), 1)
        lt, rb = self.dfl(box).chunk(2, 1)
        cos, sin = torch.cos(angle), torch.sin(angle)
        xf, yf = ((rb - lt) / 2).split(1, 1)
        xy = torch.cat((xf * cos - yf * sin, xf * sin + yf * cos), 1) + anchors.unsqueeze(0)
        dbox = torch.cat((xy, lt + rb), 1) * strides
        return torch.cat((dbox, cls.sigmoid(), angle), 1), torch.cat((raw, angle_raw), 1)


class YOLOv8(nn.Module):
    """s-scale by default (depth 0.33, width 0.50, max_channels 1024)."""

    def __init__(self, nc=4, task="detect", depth=0.33, width=0.50, max_ch=1024):
        super().__init__()
        ch = lambda c: int(math.ceil(min(c, max_ch) * width / 8) * 8)  # make_divisible(.., 8)
        rep = lambda n: max(round(n * depth), 1)
        c1, c2, c3, c4, c5 = ch(64), ch(128), ch(256), ch(512), ch(1024)
        head = OBB if task == "obb" else Detect
        self.model = nn.ModuleList([
            Conv(3, c1, 3, 2),                       # 0
            Conv(c1, c2, 3, 2),                      # 1
            C2f(c2, c2, rep(3), True),               # 2
            Conv(c2, c3, 3, 2),                      # 3
            C2f(c3, c3, rep(6), True),               # 4
            Conv(c3, c4, 3, 2),                      # 5
            C2f(c4, c4, rep(6), True),               # 6
            Conv(c4, c5, 3, 2),                      # 7
            C2f(c5, c5, rep(3), True),               # 8
            SPPF(c5, c5, 5),                         # 9
            nn.Upsample(scale_factor=2.0, mode="nearest"),  # 10
            nn.Identity(),                           # 11 concat [10, 6]
            C2f(c5 + c4, c4, rep(3)),                # 12
            nn.Upsample(scale_factor=2.0, mode="nearest"),  # 13
            nn.Identity(),                           # 14 concat [13, 4]
            C2f(c4 + c3, c3, rep(3)),                # 15 (P3)
            Conv(c3, c3, 3, 2),                      # 16
            nn.Identity(),                           # 17 concat [16, 12]
            C2f(c3 + c4, c4, rep(3)),                # 18 (P4)
            Conv(c4, c4, 3, 2),                      # 19
            nn.Identity(),                           # 20 concat [19, 9]
            C2f(c4 + c5, c5, rep(3)),                # 21 (P5)
            head(nc, (c3, c4, c5)),                  # 22
        ])
        self.nc, self.task = nc, task

    def features(self, x, taps: Dict[str, torch.Tensor] | None = None):
        m = self.model
        y = {}
        for i in range(10):
            x = m[i](x)
            y[i] = x
        x = torch.cat((m[10](y[9]), y[6]), 1)
        y[12] = m[12](x)
        x = torch.cat((m[13](y[12]), y[4]), 1)
        y[15] = m[15](x)
        x = torch.cat((m[16](y[15]), y[12]), 1)
        y[18] = m[18](x)
        x = torch.cat((m[19](y[18]), y[9]), 1)
        y[21] = m[21](x)
        if taps is not None:
            taps.update({str(k): v for k, v in y.items()})
        return [y[15], y[18], y[21]]

    def forward(self, x, taps=None):
        """Returns (decoded (B, 4+nc(+1), A), raw (B, 64+nc(+1), A))."""
        return self.model[22](self.features(x, taps))


def count_parameters(model: nn.Module) -> int:
    return sum(p.numel() for p in model.parameters())
