"""Detector weights: seeded random initialisation, BatchNorm folding, and ultralytics ``.pt`` harvesting.

State-dict keys follow the ultralytics layout (``model.<idx>...``) that the reference's checkpoint uses
(/root/reference/geotrax/extract.py:222 ``YOLO(model=config['model'])``; default weights
``geotrax_hbb_yolov8s_1920_v1.pt``, /root/reference/geotrax/cfg/default.yaml:81).  ``fold()`` reproduces
``Model.fuse()``: w' = w * gamma / sqrt(var + eps), b' = beta - mean * gamma / sqrt(var + eps), eps = 1e-3.
"""
from __future__ import annotations

import io
import math
import pickle
import zipfile
from typing import Dict, List, Tuple

import cv2
import numpy as np
import torch

BN_EPS = 1e-3


def conv_specs(nc: int = 4, task: str = "detect") -> List[Tuple[str, int, int, int, int, int]]:
    """Canonical conv list (name, cin, cout, k, stride, act) in ultralytics module order; mirrors gt_conv_info()."""
    c1, c2, c3, c4, c5 = 32, 64, 128, 256, 512
    out: List[Tuple[str, int, int, int, int, int]] = []

    def conv(name, cin, cout, k, s, act=1):
        out.append((name, cin, cout, k, s, act))

    def c2f(pre, a, b, n):
        c = b // 2
        conv(pre + ".cv1", a, 2 * c, 1, 1)
        conv(pre + ".cv2", (2 + n) * c, b, 1, 1)
        for i in range(n):
            conv(f"{pre}.m.{i}.cv1", c, c, 3, 1)
            conv(f"{pre}.m.{i}.cv2", c, c, 3, 1)

    conv("model.0", 3, c1, 3, 2); conv("model.1", c1, c2, 3, 2); c2f("model.2", c2, c2, 1)
    conv("model.3", c2, c3, 3, 2); c2f("model.4", c3, c3, 2)
    conv("model.5", c3, c4, 3, 2); c2f("model.6", c4, c4, 2)
    conv("model.7", c4, c5, 3, 2); c2f("model.8", c5, c5, 1)
    conv("model.9.cv1", c5, c5 // 2, 1, 1); conv("model.9.cv2", c5 * 2, c5, 1, 1)
    c2f("model.12", c5 + c4, c4, 1); c2f("model.15", c4 + c3, c3, 1)
    conv("model.16", c3, c3, 3, 2); c2f("model.18", c3 + c4, c4, 1)
    conv("model.19", c4, c4, 3, 2); c2f("model.21", c4 + c5, c5, 1)
    ch = (c3, c4, c5)
    h2, h3, h4 = 64, max(c3, min(nc, 100)), max(c3 // 4, 1)
    for i in range(3):
        conv(f"model.22.cv2.{i}.0", ch[i], h2, 3, 1); conv(f"model.22.cv2.{i}.1", h2, h2, 3, 1); conv(f"model.22.cv2.{i}.2", h2, 64, 1, 1, 0)
    for i in range(3):
        conv(f"model.22.cv3.{i}.0", ch[i], h3, 3, 1); conv(f"model.22.cv3.{i}.1", h3, h3, 3, 1); conv(f"model.22.cv3.{i}.2", h3, nc, 1, 1, 0)
    if task == "obb":
        for i in range(3):
            conv(f"model.22.cv4.{i}.0", ch[i], h4, 3, 1); conv(f"model.22.cv4.{i}.1", h4, h4, 3, 1); conv(f"model.22.cv4.{i}.2", h4, 1, 1, 1, 0)
    return out


def random_state_dict(nc: int = 4, task: str = "detect", seed: int = 0, cls_bias: float = -3.0, calibrate: bool = True,
                      frame_hw=(2160, 3840), imgsz: int = 1920) -> Dict[str, torch.Tensor]:
    """Unfused (Conv + BN) random-init state dict, deterministic in ``seed``; activations stay O(1) through the net."""
    g = torch.Generator().manual_seed(seed)
    sd: Dict[str, torch.Tensor] = {}
    for name, cin, cout, k, s, act in conv_specs(nc, task):
        fan_in = cin * k * k
        w = torch.randn(cout, cin, k, k, generator=g) * math.sqrt(2.0 / fan_in)
        if act:  # Conv = conv(no bias) + BN + SiLU
            sd[name + ".conv.weight"] = w
            sd[name + ".bn.weight"] = 0.8 + 0.4 * torch.rand(cout, generator=g)
            sd[name + ".bn.bias"] = 0.5 + 0.1 * torch.randn(cout, generator=g)  # positive shift keeps SiLU in its well-conditioned range
            sd[name + ".bn.running_mean"] = 0.1 * torch.randn(cout, generator=g)
            sd[name + ".bn.running_var"] = 0.8 + 0.4 * torch.rand(cout, generator=g)
        else:   # plain nn.Conv2d with bias (last conv of each head branch)
            sd[name + ".weight"] = w
            if ".cv2." in name:
                b = torch.full((cout,), 1.0)
            elif ".cv3." in name:
                b = torch.full((cout,), cls_bias) + 0.2 * torch.randn(cout, generator=g)
            else:
                b = 0.2 * torch.randn(cout, generator=g)
            sd[name + ".bias"] = b
    sd["model.22.dfl.conv.weight"] = torch.arange(16, dtype=torch.float32).view(1, 16, 1, 1)
    if calibrate:
        calibrate_bn_(sd, nc, task, seed, frame_hw, imgsz)
    return sd


def fold(sd: Dict[str, torch.Tensor], nc: int = 4, task: str = "detect") -> Dict[str, Tuple[np.ndarray, np.ndarray]]:
    """-> {conv name: (w f32 [cout,cin,k,k], b f32 [cout])} with BN folded (accepts already-fused dicts too)."""
    out = {}
    for name, cin, cout, k, s, act in conv_specs(nc, task):
        if name + ".conv.weight" in sd:
            w = sd[name + ".conv.weight"].float()
            if name + ".bn.weight" in sd:
                scale = sd[name + ".bn.weight"].float() / torch.sqrt(sd[name + ".bn.running_var"].float() + BN_EPS)
                b = sd[name + ".bn.bias"].float() - sd[name + ".bn.running_mean"].float() * scale
                w = w * scale.view(-1, 1, 1, 1)
            else:
                b = sd.get(name + ".conv.bias", torch.zeros(cout)).float()
        else:
            w = sd[name + ".weight"].float()
            b = sd[name + ".bias"].float()
        assert tuple(w.shape) == (cout, cin, k, k), f"{name}: weight {tuple(w.shape)} != {(cout, cin, k, k)}"
        out[name] = (np.ascontiguousarray(w.numpy(), dtype=np.float32), np.ascontiguousarray(b.numpy(), dtype=np.float32))
    return out


# ---- ultralytics checkpoint harvesting without ultralytics ---------------------------------------------------------------
class _Stub:
    """Inert stand-in for any pickled ultralytics/torch.nn class: keeps __dict__, builds nothing."""

    def __init__(self, *a, **k):
        pass

    def __setstate__(self, state):
        self.__dict__.update(state if isinstance(state, dict) else {"_state": state})


class _RestrictedUnpickler(pickle.Unpickler):
    """Unpickler for ultralytics checkpoints that can only rebuild tensors and plain containers.

    ``find_class`` resolves an explicit (module, name) allow-list -- tensor / storage reconstruction, ``OrderedDict``, numpy array and
    scalar reconstruction, a few harmless builtins -- and turns EVERYTHING else (the pickled ``nn.Module`` tree, but also ``eval``,
    ``os.system``, ``torch.hub.load`` ...) into an inert ``_Stub`` class whose construction runs no code.  ``copyreg._reconstructor``
    (protocol < 2 object pickles) is replaced by a version that only ever instantiates ``_Stub`` subclasses."""

    _ALLOWED = {
        ("collections", "OrderedDict"), ("collections", "defaultdict"),
        ("builtins", "set"), ("builtins", "frozenset"), ("builtins", "dict"), ("builtins", "list"), ("builtins", "tuple"), ("builtins", "int"),
        ("builtins", "float"), ("builtins", "bool"), ("builtins", "str"), ("builtins", "bytes"), ("builtins", "bytearray"), ("builtins", "complex"),
        ("builtins", "slice"), ("builtins", "range"), ("builtins", "object"),
        ("torch._utils", "_rebuild_tensor_v2"), ("torch._utils", "_rebuild_tensor"), ("torch._utils", "_rebuild_parameter"),
        ("torch._utils", "_rebuild_parameter_with_state"), ("torch._utils", "_rebuild_qtensor"), ("torch._tensor", "_rebuild_from_type_v2"),
        ("torch", "Size"), ("torch", "device"), ("torch", "dtype"), ("torch", "Tensor"), ("torch.nn.parameter", "Parameter"),
        ("torch.serialization", "_get_layout"),
        ("numpy", "ndarray"), ("numpy", "dtype"), ("numpy.core.multiarray", "_reconstruct"), ("numpy._core.multiarray", "_reconstruct"),
        ("numpy.core.multiarray", "scalar"), ("numpy._core.multiarray", "scalar"), ("_codecs", "encode"),
    }
    _TORCH_TYPED = ("Storage", "Tensor")      # torch.FloatStorage, torch.HalfStorage, torch.cuda.FloatTensor ... : data-only classes

    @staticmethod
    def _reconstructor(cls, base, state):
        """copyreg._reconstructor restricted to inert stubs (anything else is refused)."""
        if not (isinstance(cls, type) and issubclass(cls, _Stub)):
            raise pickle.UnpicklingError(f"refusing to reconstruct {cls!r} from a checkpoint")
        return object.__new__(cls)

    def find_class(self, module, name):
        if (module, name) in self._ALLOWED:
            return super().find_class(module, name)
        if module in ("torch", "torch.storage", "torch.cuda") and name.endswith(self._TORCH_TYPED) and name.isidentifier():
            obj = super().find_class(module, name)
            if isinstance(obj, type):
                return obj
        if module == "torch" and name in ("float32", "float16", "bfloat16", "float64", "int64", "int32", "int16", "int8", "uint8", "bool"):
            return getattr(torch, name)
        if (module, name) == ("copyreg", "_reconstructor"):
            return self._reconstructor
        return type(name, (_Stub,), {"__module__": module})


def _walk(obj, prefix, out, seen):
    if id(obj) in seen:
        return
    seen.add(id(obj))
    d = getattr(obj, "__dict__", None)
    if not isinstance(d, dict):
        return
    for kind in ("_parameters", "_buffers"):
        for k, v in (d.get(kind) or {}).items():
            if isinstance(v, torch.Tensor):
                out[f"{prefix}{k}"] = v.detach().float()
    for k, m in (d.get("_modules") or {}).items():
        _walk(m, f"{prefix}{k}.", out, seen)


def load_pt(path: str):
    """-> (state_dict, names, task, nc) harvested from an ultralytics checkpoint (pickled DetectionModel)."""

    class _P:
        Unpickler = _RestrictedUnpickler
        __name__ = "pickle"
        load = staticmethod(lambda f, **k: _RestrictedUnpickler(f, **k).load())

    ckpt = torch.load(path, map_location="cpu", pickle_module=_P, weights_only=False)
    model = ckpt.get("ema") or ckpt["model"] if isinstance(ckpt, dict) else ckpt
    sd: Dict[str, torch.Tensor] = {}
    _walk(model, "", sd, set())
    names = getattr(model, "names", None) or {}
    yaml_cfg = getattr(model, "yaml", {}) or {}
    nc = int(yaml_cfg.get("nc", len(names) or 4))
    task = "obb" if any(".cv4." in k for k in sd) else "detect"
    return sd, dict(names), task, nc


# ---- data-driven calibration of the random init (host-side, init time only) -------------------------------------------------
def calibrate_bn_(sd: Dict[str, torch.Tensor], nc: int = 4, task: str = "detect", seed: int = 0, frame_hw=(2160, 3840), imgsz: int = 1920) -> None:
    """Sets every BN's running_mean/var to the statistics its conv produces on a random image, as training would.

    Random conv weights otherwise make activations explode or vanish over the 20+ layers; with calibrated statistics
    every pre-activation is ~N(beta, gamma^2), which is what a trained checkpoint looks like.  In place.
    """
    import torch.nn.functional as F

    from . import synth  # calibration image has the statistics of the synthetic frames the tests / bench use

    h0, w0 = frame_hw
    img = synth.make_flight(1, h0, w0, seed + 1, n_vehicles=max(8, int(132 * w0 / 3840)))[0][0]
    r = min(imgsz / h0, imgsz / w0)
    nw, nh = int(round(w0 * r)), int(round(h0 * r))
    img = cv2.resize(img, (nw, nh), interpolation=cv2.INTER_LINEAR)
    dh, dw = ((imgsz - nh) % 32) / 2, ((imgsz - nw) % 32) / 2
    img = cv2.copyMakeBorder(img, int(round(dh - 0.1)), int(round(dh + 0.1)), int(round(dw - 0.1)), int(round(dw + 0.1)),
                             cv2.BORDER_CONSTANT, value=(114, 114, 114))
    x = torch.from_numpy(np.ascontiguousarray(img[..., ::-1].transpose(2, 0, 1))).float().unsqueeze(0) / 255.0

    def conv(name, t, k, s):
        y = F.conv2d(t, sd[name + ".conv.weight"], None, s, k // 2)
        sd[name + ".bn.running_mean"] = y.mean((0, 2, 3))
        sd[name + ".bn.running_var"] = y.var((0, 2, 3), unbiased=False)
        y = F.batch_norm(y, sd[name + ".bn.running_mean"], sd[name + ".bn.running_var"], sd[name + ".bn.weight"], sd[name + ".bn.bias"],
                         False, 0.0, BN_EPS)
        return F.silu(y)

    def c2f(pre, t, n, shortcut):
        y = list(conv(pre + ".cv1", t, 1, 1).chunk(2, 1))
        for i in range(n):
            z = conv(f"{pre}.m.{i}.cv2", conv(f"{pre}.m.{i}.cv1", y[-1], 3, 1), 3, 1)
            y.append(y[-1] + z if shortcut else z)
        return conv(pre + ".cv2", torch.cat(y, 1), 1, 1)

    up = lambda t: F.interpolate(t, scale_factor=2.0, mode="nearest")
    with torch.no_grad():
        t0 = conv("model.0", x, 3, 2); t1 = conv("model.1", t0, 3, 2); t2 = c2f("model.2", t1, 1, True)
        t3 = conv("model.3", t2, 3, 2); t4 = c2f("model.4", t3, 2, True)
        t5 = conv("model.5", t4, 3, 2); t6 = c2f("model.6", t5, 2, True)
        t7 = conv("model.7", t6, 3, 2); t8 = c2f("model.8", t7, 1, True)
        a = conv("model.9.cv1", t8, 1, 1)
        p = [a]
        for _ in range(3):
            p.append(F.max_pool2d(p[-1], 5, 1, 2))
        t9 = conv("model.9.cv2", torch.cat(p, 1), 1, 1)
        t12 = c2f("model.12", torch.cat((up(t9), t6), 1), 1, False)
        t15 = c2f("model.15", torch.cat((up(t12), t4), 1), 1, False)
        t18 = c2f("model.18", torch.cat((conv("model.16", t15, 3, 2), t12), 1), 1, False)
        t21 = c2f("model.21", torch.cat((conv("model.19", t18, 3, 2), t9), 1), 1, False)
        for i, f in enumerate((t15, t18, t21)):
            for br, target in ((("cv2", 1.5), ("cv3", 1.0), ("cv4", 1.0)) if task == "obb" else (("cv2", 1.5), ("cv3", 1.0))):
                z = conv(f"model.22.{br}.{i}.1", conv(f"model.22.{br}.{i}.0", f, 3, 1), 3, 1)
                wk = f"model.22.{br}.{i}.2.weight"
                sd[wk] = sd[wk] * (target / float(F.conv2d(z, sd[wk]).std() + 1e-6))  # logits of O(1) spread
