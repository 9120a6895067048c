"""Parity at BASELINE.json's full size (3840x2160 frames, imgsz 1920 -> 1088x1920, 42,840 anchors, ORB 2000/4000): the oracle where it
finishes in seconds (one fp32 CPU forward, decode + NMS, OpenCV ORB / MAGSAC on one frame pair) and size-independent properties
elsewhere (idempotence, sortedness, NMS separation, box-warp semantics, reference frame -> identity)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

HW, IMGSZ = (2160, 3840), 1920


@pytest.fixture(scope="module")
def full_engine():
    import geotrax_b200
    from geotrax_b200 import weights
    eng = geotrax_b200.Engine(frame_hw=HW, imgsz=IMGSZ, nc=4, max_batch=2)
    sd = weights.random_state_dict(4, "detect", seed=0, frame_hw=HW, imgsz=IMGSZ, cls_bias=-4.4)
    eng.load_weights(weights.fold(sd))
    eng._sd = sd
    yield eng
    eng.close()


@pytest.fixture(scope="module")
def full_flight():
    from geotrax_b200 import synth
    return synth.make_flight(3, HW[0], HW[1], seed=100)


def test_fullsize_preprocess_bit_exact(full_engine):
    import cv2
    from oracle import prepost
    rng = np.random.default_rng(5)
    frames = rng.integers(0, 256, (2,) + HW + (3,), dtype=np.uint8)
    full_engine.preprocess(frames)
    got = full_engine.net_input(2)
    ref = prepost.preprocess(list(frames), IMGSZ).numpy()
    assert got.shape == ref.shape == (2, 3, 1088, 1920)
    assert np.array_equal(got.astype(np.float32) / np.float32(255.0), ref)
    gray = full_engine.gray(2)
    for i in range(2):
        g = cv2.resize(cv2.cvtColor(frames[i], cv2.COLOR_BGR2GRAY), (1920, 1080), interpolation=cv2.INTER_LINEAR)
        assert np.array_equal(gray[i], g)


def test_fullsize_raw_head_and_detections(full_engine, full_flight):
    """criterion (1) at the real size on one frame (a 145-GFLOP fp32 CPU forward), criterion (2) on the same raw head, plus the
    size-independent NMS properties on both frames."""
    from oracle import prepost
    from oracle.yolov8 import YOLOv8
    eng = full_engine
    frames = np.stack(full_flight[0][1:3])
    eng.preprocess(frames)
    boxes, counts, keep = eng.detect(2, conf=0.25, iou=0.7, agnostic=True, classes=[0, 1, 2, 3], want_keep=True)
    boxes2, counts2, keep2 = eng.detect(2, conf=0.25, iou=0.7, agnostic=True, classes=[0, 1, 2, 3], want_keep=True)
    assert np.array_equal(boxes, boxes2) and np.array_equal(counts, counts2) and np.array_equal(keep, keep2), "detect is not idempotent"
    raw = eng.raw_head(2)
    assert raw.shape == (2, 42840, 68)
    m = YOLOv8(4).eval()
    m.load_state_dict(eng._sd, strict=False)
    with torch.no_grad():
        _, ref = m(prepost.preprocess([frames[0]], IMGSZ))
    ref = ref.permute(0, 2, 1).numpy()[0]
    rel = np.linalg.norm(raw[0] - ref) / np.linalg.norm(ref)
    print("full-size raw head rel L2 error", rel)
    assert rel < 1e-2
    # decode + NMS of the oracle on the GPU's own raw head: identical keep indices, boxes within float tolerance
    head = m.model[22]
    dec = head.decode(torch.from_numpy(raw).permute(0, 2, 1).contiguous(), [(136, 240), (68, 120), (34, 60)])
    outs, idxs = prepost.non_max_suppression(dec, 0.25, 0.7, [0, 1, 2, 3], True, 1000, nc=4, return_idxs=True)
    assert counts.sum() > 50
    for b in range(2):
        n = int(counts[b])
        r = boxes[b, :n]
        want = outs[b].clone()
        want[:, :4] = prepost.scale_boxes((1088, 1920), want[:, :4], HW)
        got_set, ref_set = set(keep[b, :n].tolist()), set(idxs[b].numpy().tolist())
        assert len(got_set ^ ref_set) <= max(1, len(ref_set) // 50), f"keep sets differ by {len(got_set ^ ref_set)} of {len(ref_set)}"
        # properties that hold at any size
        assert n <= 1000 and np.all(np.diff(r[:, 4]) <= 0), "confidences must be sorted descending"
        assert np.all(r[:, 4] > 0.25) and set(np.unique(r[:, 5]).astype(int)) <= {0, 1, 2, 3}
        assert r[:, 0].min() >= 0 and r[:, 1].min() >= 0 and r[:, 2].max() <= HW[1] and r[:, 3].max() <= HW[0]
        assert len(got_set) == n, "an anchor was kept twice"
        # greedy NMS leaves no pair above the IoU threshold (checked in network coordinates, where the suppression ran)
        k = torch.from_numpy(keep[b, :n].astype(np.int64))
        nb = prepost.xywh2xyxy(dec[b, :4, k].T)
        x1, y1 = torch.maximum(nb[:, None, 0], nb[None, :, 0]), torch.maximum(nb[:, None, 1], nb[None, :, 1])
        x2, y2 = torch.minimum(nb[:, None, 2], nb[None, :, 2]), torch.minimum(nb[:, None, 3], nb[None, :, 3])
        inter = (x2 - x1).clamp(min=0) * (y2 - y1).clamp(min=0)
        area = (nb[:, 2] - nb[:, 0]) * (nb[:, 3] - nb[:, 1])
        iou = inter / (area[:, None] + area[None, :] - inter)
        iou.fill_diagonal_(0)
        assert float(iou.max()) <= 0.7 + 1e-5


def test_fullsize_stabilize(full_engine, full_flight):
    """criterion (3) at the real size on one frame pair against the OpenCV oracle and the generator's ground truth; the reference frame
    maps to the identity; the box warp is the envelope of the four warped corners."""
    from oracle.stabilo_cv import Stabilizer, warp_boxes_xywh
    eng = full_engine
    frames, boxes, Hs = full_flight
    eng.preprocess(np.stack(frames[:1]))
    eng.set_reference(0, boxes[0])
    eng.preprocess(np.stack([frames[0], frames[1]]))
    H, status, stats = eng.stabilize(2, [boxes[0], boxes[1]])
    H2, status2, stats2 = eng.stabilize(2, [boxes[0], boxes[1]])
    assert np.array_equal(H, H2) and np.array_equal(stats, stats2), "stabilize is not idempotent"
    assert status.tolist() == [0, 0]
    assert np.allclose(H[0], np.eye(3), atol=1e-5), "the reference frame must map to the identity"
    assert stats[1][0] <= 4000 + 8 * 64 and 1800 <= stats[1][1] <= 2000 + 8 * 64 and stats[1][2] >= 1000 and stats[1][3] >= 500, stats[1]
    ora = Stabilizer()
    ora.set_ref_frame(frames[0], boxes[0])
    ora.stabilize(frames[1], boxes[1])
    got = eng.warp_boxes(H[1], boxes[1])
    want, truth = ora.transform_cur_boxes(), warp_boxes_xywh(boxes[1], Hs[1])
    assert np.linalg.norm(got[:, :2] - want[:, :2], axis=1).mean() < 0.5
    assert np.linalg.norm(got[:, :2] - truth[:, :2], axis=1).mean() < 0.5
    assert np.allclose(got, warp_boxes_xywh(boxes[1], H[1]), atol=2e-3), "box warp is not the envelope of the warped corners"
    assert abs(H[1][2, 2] - 1.0) < 1e-12 and np.linalg.det(H[1]) > 0
