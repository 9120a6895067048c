"""Pins the oracle's YOLOv8s restatement and pre/post-processing against published facts and the engines the reference runs
(torch / torchvision / OpenCV), on the CPU.  SURVEY.md section 8c, Appendix A-1/A-2, B-4/B-5/B-8."""
import math

import cv2
import numpy as np
import pytest
import torch
import torchvision

from oracle import prepost
from oracle.yolov8 import YOLOv8, count_parameters


@pytest.mark.parametrize("nc,task,params", [(80, "detect", 11_166_560), (4, "detect", 11_137_148), (6, "detect", 11_137_922), (4, "obb", 11_423_327)])
def test_parameter_counts_match_published(nc, task, params):
    # 11,166,560 @ nc=80 is the published YOLOv8s figure; the others were verified in the survey (Appendix A-1)
    m = YOLOv8(nc, task)
    n = count_parameters(m)   # includes the 16 frozen DFL weights, as the published figure does
    assert n == params


def test_output_shapes_and_anchor_count():
    m = YOLOv8(4).eval()
    with torch.no_grad():
        dec, raw = m(torch.zeros(1, 3, 64, 96))
    A = 8 * 12 + 4 * 6 + 2 * 3
    assert dec.shape == (1, 8, A) and raw.shape == (1, 68, A)
    # default preset: 1088 x 1920 -> 42,840 anchors
    assert 136 * 240 + 68 * 120 + 34 * 60 == 42840


def test_letterbox_default_preset_geometry():
    new_w, new_h, top, bottom, left, right, r = prepost.letterbox_params(2160, 3840, 1920)
    assert (new_w, new_h, top, bottom, left, right, r) == (1920, 1080, 4, 4, 0, 0, 0.5)


def test_half_resize_and_gray_integer_formulas():
    """cv2.resize at exactly 1/2 == (a+b+c+d+2)>>2 and BGR2GRAY == 15-bit fixed point: what the CUDA preprocess computes."""
    rng = np.random.default_rng(0)
    img = rng.integers(0, 256, (64, 96, 3), dtype=np.uint8)
    half = cv2.resize(img, (48, 32), interpolation=cv2.INTER_LINEAR)
    s = img.astype(np.uint32)
    mine = (s[0::2, 0::2] + s[0::2, 1::2] + s[1::2, 0::2] + s[1::2, 1::2] + 2) >> 2
    assert np.array_equal(half, mine.astype(np.uint8))
    g = cv2.cvtColor(img, cv2.COLOR_BGR2GRAY)
    mine_g = (9798 * s[..., 2] + 19235 * s[..., 1] + 3735 * s[..., 0] + 16384) >> 15
    assert np.array_equal(g, mine_g.astype(np.uint8))


def test_preprocess_layout():
    rng = np.random.default_rng(1)
    f = rng.integers(0, 256, (512, 768, 3), dtype=np.uint8)
    x = prepost.preprocess([f], 384)
    assert x.shape == (1, 3, 256, 384) and x.dtype == torch.float32
    half = cv2.resize(f, (384, 256), interpolation=cv2.INTER_LINEAR)
    assert torch.equal(x[0, 0], torch.from_numpy(half[..., 2].astype(np.float32)) / 255.0)   # channel 0 is R


def test_nms_semantics_match_torchvision():
    # strict '>' on IoU, stable tie order, class offset == batched_nms
    b = torch.tensor([[0, 0, 10, 10], [0, 0, 10, 10.0], [0, 5, 10, 15], [100, 100, 110, 110]], dtype=torch.float32)
    s = torch.tensor([0.9, 0.9, 0.8, 0.7])
    assert torchvision.ops.nms(b, s, 0.5).tolist() == [0, 2, 3]       # IoU(0,2) = 1/3 survives; duplicate 1 suppressed; lower index first
    pred = torch.zeros(1, 8, 4)
    pred[0, :4] = torch.tensor([[5, 5, 10, 10], [5, 5, 10, 10], [5, 10, 10, 10], [105, 105, 10, 10]]).t()
    pred[0, 4, 0], pred[0, 5, 1], pred[0, 4, 2], pred[0, 6, 3] = 0.9, 0.9, 0.8, 0.7
    out_agn, idx_agn = prepost.non_max_suppression(pred, 0.25, 0.5, None, True, 300, nc=4, return_idxs=True)
    out_cls, idx_cls = prepost.non_max_suppression(pred, 0.25, 0.5, None, False, 300, nc=4, return_idxs=True)
    assert idx_agn[0].tolist() == [0, 2, 3]
    assert idx_cls[0].tolist() == [0, 1, 2, 3]                        # class-aware: the class-1 duplicate survives
    assert out_agn[0][:, 5].tolist() == [0.0, 0.0, 2.0]


def test_scale_boxes_default_preset():
    b = torch.tensor([[100.0, 104.0, 200.0, 204.0], [-5.0, 0.0, 1925.0, 1090.0]])
    out = prepost.scale_boxes((1088, 1920), b, (2160, 3840))
    assert torch.allclose(out[0], torch.tensor([200.0, 200.0, 400.0, 400.0]))
    assert out[1].tolist() == [0.0, 0.0, 3840.0, 2160.0]            # clipped to the frame


def test_probiou_properties():
    a = torch.tensor([[50.0, 50.0, 40.0, 20.0, 0.3]])
    assert prepost.batch_probiou(a, a).item() > 0.99
    far = torch.tensor([[500.0, 500.0, 40.0, 20.0, 0.3]])
    assert prepost.batch_probiou(a, far).item() < 1e-3
    b = torch.tensor([[55.0, 52.0, 30.0, 25.0, 1.0]])
    assert math.isclose(prepost.batch_probiou(a, b).item(), prepost.batch_probiou(b, a).item(), rel_tol=1e-5)


def test_rotated_nms_is_fast_nms():
    """A suppressed box still suppresses (Fast-NMS): chain a > b > c with iou(a,b), iou(b,c) high and iou(a,c) low."""
    boxes = torch.tensor([[0.0, 0, 40, 10, 0], [12.0, 0, 40, 10, 0], [24.0, 0, 40, 10, 0]])
    scores = torch.tensor([0.9, 0.8, 0.7])
    iou = prepost.batch_probiou(boxes, boxes)
    thr = float((iou[0, 1] + iou[0, 2]) / 2)
    assert iou[0, 1] > thr > iou[0, 2] and iou[1, 2] > thr
    keep = prepost.nms_rotated(boxes, scores, thr)
    assert keep.tolist() == [0]          # greedy NMS would keep [0, 2]


def test_regularize_rboxes():
    rb = torch.tensor([[0.0, 0, 10, 20, math.pi * 0.75], [0.0, 0, 10, 20, 0.1]])
    out = prepost.regularize_rboxes(rb)
    assert torch.allclose(out[0], torch.tensor([0.0, 0, 20, 10, math.pi * 0.25]), atol=1e-6)
    assert torch.allclose(out[1], rb[1])


def test_oracle_stabilizer_recovers_known_homography():
    """The OpenCV-backed stabilizer restatement finds the synthetic flight's ground-truth H (criterion-3 scale: < 0.5 px)."""
    import importlib.util
    import os
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    import geotrax_b200.synth as synth   # pure numpy/cv2 generator; importing it does not load the CUDA library
    from oracle.stabilo_cv import Stabilizer, warp_boxes_xywh
    frames, boxes, Hs = synth.make_flight(2, 540, 960, seed=5, n_vehicles=16)
    st = Stabilizer(max_features=1500)
    st.set_ref_frame(frames[0], boxes[0])
    st.stabilize(frames[1], boxes[1])
    H = st.get_cur_trans_matrix()
    assert H is not None and abs(H[2, 2] - 1) < 1e-12
    got = st.transform_cur_boxes()
    want = warp_boxes_xywh(boxes[1], Hs[1])
    assert np.linalg.norm(got[:, :2] - want[:, :2], axis=1).mean() < 0.5
    nr, ncur = st.get_cur_num_keypoints()
    assert nr > ncur > 500 and st.get_cur_num_matches() >= st.get_cur_inliers_count() > 100


@pytest.mark.parametrize("shape", [((1520, 2704), (1079, 1920)), ((1080, 1920), (960, 1706)), ((720, 1280), (1080, 1920)), ((333, 517), (258, 400)),
                                   ((2160, 3840), (1080, 1920)), ((480, 640), (480, 640))])
@pytest.mark.parametrize("cn", [1, 3])
def test_resize_linear_restatement_matches_cv2(shape, cn):
    """oracle/prepost.py:resize_linear_u8 (the integer formulas the CUDA general letterbox follows) is bit-exact against cv2.resize."""
    import cv2
    from oracle import prepost
    (h, w), (nh, nw) = shape
    rng = np.random.default_rng(h * 7 + cn)
    img = rng.integers(0, 256, (h, w, 3) if cn == 3 else (h, w), dtype=np.uint8)
    ref = cv2.resize(img, (nw, nh), interpolation=cv2.INTER_LINEAR)
    assert np.array_equal(prepost.resize_linear_u8(img, nw, nh), ref)


@pytest.mark.parametrize("hw", [(64, 96), (270, 480), (1080, 1920)])
def test_nv12_restatement_matches_cv2(hw):
    """oracle/prepost.py:nv12_to_bgr (the arithmetic of the CUDA decoder-format ingest) is bit-exact against cv2.cvtColor."""
    import cv2
    from oracle import prepost
    rng = np.random.default_rng(hw[0])
    nv12 = rng.integers(0, 256, (hw[0] * 3 // 2, hw[1]), dtype=np.uint8)
    assert np.array_equal(prepost.nv12_to_bgr(nv12), cv2.cvtColor(nv12, cv2.COLOR_YUV2BGR_NV12))
    frame = rng.integers(0, 256, hw + (3,), dtype=np.uint8)
    rt = prepost.bgr_to_nv12(frame)                                    # helper layout check: Y plane + interleaved chroma
    assert rt.shape == (hw[0] * 3 // 2, hw[1]) and np.array_equal(rt[:hw[0]], cv2.cvtColor(frame, cv2.COLOR_BGR2YUV_I420)[:hw[0]])


@pytest.mark.parametrize("shape", [(270, 480), (135, 250), (547, 961)])
def test_clahe_restatement_matches_cv2(shape):
    """oracle/prepost.py:clahe_u8 == cv2.createCLAHE(2.0, (8, 8)).apply, bit for bit (divisible and reflect-extended tile grids)."""
    import cv2
    from oracle import prepost
    rng = np.random.default_rng(shape[0])
    base = cv2.GaussianBlur(rng.integers(0, 256, shape, dtype=np.uint8), (0, 0), 3)
    img = np.clip(base.astype(np.int32) + rng.integers(-25, 25, shape), 0, 255).astype(np.uint8)
    img[: shape[0] // 3, : shape[1] // 4] //= 4          # a dark, low-contrast corner: clipping and redistribution actually happen
    assert np.array_equal(prepost.clahe_u8(img), cv2.createCLAHE(clipLimit=2.0, tileGridSize=(8, 8)).apply(img))


def test_warp_perspective_restatement_matches_cv2():
    """oracle/prepost.py:warp_perspective_u8 == cv2.warpPerspective (INTER_LINEAR, constant border), bit for bit, colour and gray."""
    from oracle import prepost
    rng = np.random.default_rng(4)
    for shape in ((135, 250, 3), (270, 480)):
        src = rng.integers(0, 256, shape, dtype=np.uint8)
        a, sc = np.deg2rad(rng.uniform(-3, 3)), 1 + rng.uniform(-0.03, 0.03)
        H = np.array([[sc * np.cos(a), -sc * np.sin(a), rng.uniform(-15, 15)], [sc * np.sin(a), sc * np.cos(a), rng.uniform(-15, 15)],
                      [rng.uniform(-2e-5, 2e-5), rng.uniform(-2e-5, 2e-5), 1.0]])
        assert np.array_equal(prepost.warp_perspective_u8(src, H), cv2.warpPerspective(src, H, (shape[1], shape[0])))
