import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu on the GPU box")


def f32_to_bf16_bits(x: np.ndarray) -> np.ndarray:
    """float32 -> bf16 bit pattern (uint16), round-to-nearest-even."""
    u = np.ascontiguousarray(x, np.float32).view(np.uint32).astype(np.uint64)
    r = ((u >> 16) & 1) + 0x7FFF
    return ((u + r) >> 16).astype(np.uint16)


def bf16_bits_to_f32(b: np.ndarray) -> np.ndarray:
    return (b.astype(np.uint32) << 16).view(np.float32)


def bf16_round(x: np.ndarray) -> np.ndarray:
    return bf16_bits_to_f32(f32_to_bf16_bits(x))


def _make_engine(act_dtype, frame_hw=(512, 768), imgsz=384, **kw):
    import geotrax_b200
    from geotrax_b200 import weights

    eng = geotrax_b200.Engine(frame_hw=frame_hw, imgsz=imgsz, nc=4, max_batch=2, max_det=300, max_features=500, act_dtype=act_dtype, **kw)
    sd = weights.random_state_dict(4, "detect", seed=0, frame_hw=frame_hw, imgsz=imgsz, cls_bias=-4.0)
    eng.load_weights(weights.fold(sd))
    eng._sd = sd
    return eng


@pytest.fixture(scope="session")
def small_engine():
    """Engine on a 512x768 frame (imgsz 384 -> net 256x384), batch 2, seeded random weights, fp16 activations (default)."""
    eng = _make_engine("fp16")
    yield eng
    eng.close()


@pytest.fixture(scope="session")
def small_engine_tc():
    """Same, forced onto the pixel-major conv kernel (conv_tc.cu); the default engines autotune and unit-test the swapped one."""
    import os
    old = os.environ.get("GT_SWAP")
    os.environ["GT_SWAP"] = "0"
    try:
        eng = _make_engine("fp16")
    finally:
        if old is None:
            os.environ.pop("GT_SWAP")
        else:
            os.environ["GT_SWAP"] = old
    yield eng
    eng.close()


def _forced_engine(swap):
    import os
    old = os.environ.get("GT_SWAP")
    os.environ["GT_SWAP"] = str(swap)
    try:
        return _make_engine("fp16")
    finally:
        if old is None:
            os.environ.pop("GT_SWAP")
        else:
            os.environ["GT_SWAP"] = old


@pytest.fixture(scope="session")
def small_engine_sw():
    """Forced onto the swapped-operand kernel without halo staging (GT_SWAP=1)."""
    eng = _forced_engine(1)
    yield eng
    eng.close()


@pytest.fixture(scope="session")
def small_engine_occ2():
    """Forced onto the pixel-major kernel at two CTAs per SM wherever it applies (GT_SWAP=3)."""
    eng = _forced_engine(3)
    yield eng
    eng.close()


@pytest.fixture(scope="session")
def small_engine_tc_halo():
    """Forced onto the pixel-major kernel with halo staging wherever it applies (GT_SWAP=4)."""
    eng = _forced_engine(4)
    yield eng
    eng.close()


@pytest.fixture(scope="session")
def small_engine_occ2_halo():
    """Two CTAs per SM + halo staging wherever they apply (GT_SWAP=5)."""
    eng = _forced_engine(5)
    yield eng
    eng.close()


@pytest.fixture(scope="session")
def small_engine_pair():
    """Swapped kernel as CTA pairs (cta_group::2) on every layer whose cout is a multiple of 256 (GT_SWAP=6)."""
    eng = _forced_engine(6)
    yield eng
    eng.close()


@pytest.fixture(scope="session")
def small_engine_bf16():
    eng = _make_engine("bf16")
    yield eng
    eng.close()


@pytest.fixture(scope="session")
def mid_engine():
    """1024x1536 frame (imgsz 768 -> net 512x768, 8064 anchors): big enough for the > 4096-candidate NMS path."""
    eng = _make_engine("fp16", frame_hw=(1024, 1536), imgsz=768)
    yield eng
    eng.close()
