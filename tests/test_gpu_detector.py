"""GPU parity of stage 1 + stage 2 against the oracle, through the C-ABI (geo-trax_b200/_lib.py -> libgeotrax_b200.so)."""
import numpy as np
import pytest
import torch

from conftest import bf16_bits_to_f32, bf16_round, f32_to_bf16_bits

pytestmark = pytest.mark.gpu


def _frames(n, h, w, seed=0):
    from geotrax_b200 import synth
    return np.stack(synth.make_flight(n, h, w, seed, n_vehicles=20)[0])


def _oracle_model(sd, nc=4, task="detect"):
    from oracle.yolov8 import YOLOv8
    m = YOLOv8(nc, task).eval()
    missing = m.load_state_dict(sd, strict=False)
    assert not missing.unexpected_keys
    return m


def test_preprocess_bit_exact(small_engine):
    import cv2
    from oracle import prepost
    eng = small_engine
    rng = np.random.default_rng(0)
    frames = rng.integers(0, 256, (2, 512, 768, 3), dtype=np.uint8)
    eng.preprocess(frames)
    got = eng.net_input(2)                      # u8 planar RGB; the 1/255 is folded into layer 0's weights
    ref = prepost.preprocess(list(frames), 384).numpy()
    assert got.shape == ref.shape == (2, 3, 256, 384)
    assert np.array_equal(got.astype(np.float32) / np.float32(255.0), ref), "letterbox differs from the oracle (bit-exact expected)"
    gray = eng.gray(2)
    for i in range(2):
        g = cv2.resize(cv2.cvtColor(frames[i], cv2.COLOR_BGR2GRAY), (384, 256), interpolation=cv2.INTER_LINEAR)
        assert np.array_equal(gray[i], g), "gray half-res differs from cv2 (bit-exact expected)"


@pytest.mark.parametrize("hw,imgsz,ratio", [((380, 676), 480, 0.5),      # 0.71x letterbox (2704x1520-like), even frame
                                            ((270, 480), 480, 1.0),      # identity letterbox, full-resolution working image
                                            ((375, 667), 384, 0.5),      # odd frame: bilinear working image too
                                            ((512, 360), 256, 0.5)])     # portrait: exact 1/2 but rows not 16-pixel aligned (2x2 decimation path)
def test_preprocess_general_geometry_bit_exact(hw, imgsz, ratio):
    """Any frame size / imgsz / downsample_ratio: letterbox and gray working image bit-exact against cv2 (table-driven kernels)."""
    import cv2
    import geotrax_b200
    from oracle import prepost
    eng = geotrax_b200.Engine(frame_hw=hw, imgsz=imgsz, nc=4, max_batch=2, max_det=100, max_features=300, downsample_ratio=ratio)
    try:
        rng = np.random.default_rng(1)
        frames = rng.integers(0, 256, (2,) + hw + (3,), dtype=np.uint8)
        eng.preprocess(frames)
        got = eng.net_input(2)
        ref = prepost.preprocess(list(frames), imgsz).numpy()
        assert got.shape == ref.shape
        assert np.array_equal(got.astype(np.float32) / np.float32(255.0), ref), "general letterbox differs from cv2 / ultralytics LetterBox"
        gray = eng.gray(2)
        ww, wh = int(hw[1] * ratio), int(hw[0] * ratio)
        for i in range(2):
            g = cv2.cvtColor(frames[i], cv2.COLOR_BGR2GRAY)
            if ratio != 1.0:
                g = cv2.resize(g, (ww, wh), interpolation=cv2.INTER_LINEAR)
            assert np.array_equal(gray[i], g), "gray working image differs from cv2"
    finally:
        eng.close()


@pytest.mark.parametrize("hw,imgsz", [((512, 768), 384), ((380, 676), 480)])
def test_nv12_ingest_bit_exact(hw, imgsz):
    """Decoder-format ingest (gt_set_input_format(GT_INPUT_NV12)): network input and gray working image equal the BGR path fed with
    cv2.cvtColor(COLOR_YUV2BGR_NV12) of the same NV12 frames -- bit for bit, for the fused default kernel and the general kernels."""
    import cv2
    import geotrax_b200
    from oracle import prepost
    eng = geotrax_b200.Engine(frame_hw=hw, imgsz=imgsz, nc=4, max_batch=2, max_det=100, max_features=300)
    try:
        rng = np.random.default_rng(3)
        nv12 = rng.integers(0, 256, (2, hw[0] * 3 // 2, hw[1]), dtype=np.uint8)
        bgr = np.stack([cv2.cvtColor(f, cv2.COLOR_YUV2BGR_NV12) for f in nv12])
        assert np.array_equal(bgr[0], prepost.nv12_to_bgr(nv12[0]))
        eng.set_input_format("nv12")
        eng.preprocess(nv12)
        got, gray = eng.net_input(2), eng.gray(2)
        ref = prepost.preprocess(list(bgr), imgsz).numpy()
        assert np.array_equal(got.astype(np.float32) / np.float32(255.0), ref)
        eng.set_input_format("bgr24")
        eng.preprocess(bgr)
        assert np.array_equal(eng.net_input(2), got) and np.array_equal(eng.gray(2), gray)
    finally:
        eng.close()


CONV_CASES = [
    # (B, H, W, cin, cout, k, stride, act, residual, f32)
    (2, 32, 48, 64, 64, 1, 1, True, False, False),
    (2, 32, 48, 64, 128, 3, 1, True, False, False),
    (1, 34, 60, 32, 32, 3, 1, True, True, False),      # cin < 64 (zero-filled K), residual, ragged tiles
    (2, 32, 48, 64, 128, 3, 2, True, False, False),     # stride 2 (TMA elementStrides)
    (1, 17, 30, 96, 64, 1, 1, True, False, False),      # cin not a multiple of 64
    (1, 16, 24, 128, 512, 1, 1, True, False, False),    # two N tiles
    (1, 16, 24, 256, 192, 3, 1, True, False, False),    # fused head width (N = 192)
    (1, 20, 28, 128, 4, 1, 1, False, False, True),      # final class conv: N padded to 16, fp32 rows
    (1, 20, 28, 64, 64, 1, 1, False, False, True),      # final box conv, fp32 rows
    (1, 68, 120, 128, 128, 3, 2, True, False, False),
    (2, 40, 44, 64, 64, 3, 1, True, True, False),       # halo staging with the weight ring, ragged 8 x 32 tiles, residual
    (1, 70, 20, 128, 64, 3, 1, True, False, False),     # halo staging, two k-blocks per tap
    (1, 34, 60, 128, 256, 3, 1, True, True, False),     # CTA-pair tile (cout % 256 == 0): ragged tiles, residual
    (2, 68, 120, 128, 256, 3, 2, True, False, False),   # CTA-pair tile, stride 2 (half-height boxes with elementStrides)
    (16, 34, 60, 256, 512, 1, 1, True, False, False),   # more pair tiles than TPCs, two pair tiles along cout
]


@pytest.mark.parametrize("dtype", ["fp16", "bf16", "fp16-pixel-major", "fp16-swapped-nohalo", "fp16-two-cta", "fp16-tc-halo", "fp16-two-cta-halo", "fp16-pair"])
@pytest.mark.parametrize("case", CONV_CASES)
def test_conv2d_tcgen05_matches_torch(request, case, dtype):
    fixture = {"fp16": "small_engine", "bf16": "small_engine_bf16", "fp16-pixel-major": "small_engine_tc", "fp16-swapped-nohalo": "small_engine_sw",
               "fp16-two-cta": "small_engine_occ2", "fp16-tc-halo": "small_engine_tc_halo", "fp16-two-cta-halo": "small_engine_occ2_halo",
               "fp16-pair": "small_engine_pair"}[dtype]
    eng = request.getfixturevalue(fixture)
    rnd = lambda a: eng.act_to_f32(eng.f32_to_act(a))
    B, H, W, cin, cout, k, s, act, use_res, f32 = case
    g = torch.Generator().manual_seed(abs(hash(case)) % (2 ** 31))
    x = rnd(torch.randn(B, H, W, cin, generator=g).numpy())
    w = (torch.randn(cout, cin, k, k, generator=g) / (cin * k * k) ** 0.5).numpy()
    b = torch.randn(cout, generator=g).numpy() * 0.1
    ref = torch.nn.functional.conv2d(torch.from_numpy(x).permute(0, 3, 1, 2), torch.from_numpy(rnd(w)), torch.from_numpy(b), s, k // 2)
    if act:
        ref = torch.nn.functional.silu(ref)
    res_bits = None
    if use_res:
        res = rnd(torch.randn(*ref.permute(0, 2, 3, 1).shape, generator=g).numpy())
        ref = ref + torch.from_numpy(res).permute(0, 3, 1, 2)
        res_bits = eng.f32_to_act(res)
    ref = ref.permute(0, 2, 3, 1).numpy()
    out = eng.conv2d(eng.f32_to_act(x), w, b, k, s, act, res_bits, out_f32=f32)
    got = out if f32 else eng.act_to_f32(out)
    assert got.shape == ref.shape
    # fp32 rows: accumulation order only; 16-bit rows: one output rounding on top (2^-9 bf16, 2^-12 fp16)
    tol = 2e-3 if (f32 or dtype != "bf16") else 1e-2
    err = np.abs(got - ref).max() / max(1e-6, np.abs(ref).max())
    assert err < tol, f"conv {case}: max-normalised error {err:.3e}"


def _emulated_16bit_error(sd, x, dt):
    """CPU emulation: fp32 oracle with every Conv output rounded to the 16-bit format (what storing activations costs)."""
    from oracle.yolov8 import Conv
    m = _oracle_model(sd)
    hooks = [mod.register_forward_hook(lambda mod, i, o: o.to(dt).float()) for mod in m.modules() if isinstance(mod, Conv)]
    with torch.no_grad():
        _, r = m(x)
    for h in hooks:
        h.remove()
    return r


@pytest.mark.parametrize("dtype", ["fp16", "bf16"])
def test_raw_head_within_1e2_of_fp32_oracle(small_engine, small_engine_bf16, dtype):
    """north_star criterion (1): raw head tensor within 1e-2 relative (L2) of the fp32 reference.

    fp16 storage (the default) meets 1e-2.  bf16 storage is a pure format-precision question on these random-init
    weights: rounding every activation to bf16 in the fp32 CPU oracle gives the same few-percent error, so the bf16 mode is
    held to 2x that emulated error instead (and to 1e-2 per layer in test_conv2d_*)."""
    from oracle import prepost
    eng = small_engine if dtype == "fp16" else small_engine_bf16
    frames = _frames(2, 512, 768, seed=3)
    eng.preprocess(frames)
    eng.detect(2, conf=0.25)
    raw = eng.raw_head(2)                                  # (B, A, no)
    m = _oracle_model(eng._sd)
    with torch.no_grad():
        dec, ref = m(prepost.preprocess(list(frames), 384))  # (B, no, A)
    ref = ref.permute(0, 2, 1).numpy()
    assert raw.shape == ref.shape
    rel = np.linalg.norm(raw - ref) / np.linalg.norm(ref)
    print(dtype, "raw head rel L2 error", rel, "max abs", np.abs(raw - ref).max(), "ref absmax", np.abs(ref).max())
    if dtype == "fp16":
        assert rel < 1e-2
    else:
        emu = _emulated_16bit_error(eng._sd, prepost.preprocess(list(frames), 384), torch.bfloat16).permute(0, 2, 1).numpy()
        rel_emu = np.linalg.norm(emu - ref) / np.linalg.norm(ref)
        print("bf16 emulated-on-CPU rel L2 error", rel_emu)
        assert rel < max(2.5 * rel_emu, 1e-2)   # the GPU also rounds the folded weights to bf16


@pytest.mark.parametrize("variant", ["pixel-major", "two-cta", "swapped-nohalo", "tc-halo", "two-cta-halo", "pair"])
def test_forced_kernel_raw_head(request, variant):
    """Each forced conv kernel variant (conv_tc.cu at one / two CTAs per SM, conv_sw.cu without halo) through the whole network."""
    from oracle import prepost
    eng = request.getfixturevalue({"pixel-major": "small_engine_tc", "two-cta": "small_engine_occ2", "swapped-nohalo": "small_engine_sw",
                                   "tc-halo": "small_engine_tc_halo", "two-cta-halo": "small_engine_occ2_halo", "pair": "small_engine_pair"}[variant])
    if variant == "pair":
        assert eng.conv_pair_count() >= 10, "the cout % 256 == 0 layers must run on the CTA-pair kernel"
    elif variant != "swapped-nohalo":
        assert eng.conv_kernel_info()[1] == 0
    frames = _frames(2, 512, 768, seed=3)
    eng.preprocess(frames)
    eng.detect(2, conf=0.25)
    raw = eng.raw_head(2)
    m = _oracle_model(eng._sd)
    with torch.no_grad():
        _, ref = m(prepost.preprocess(list(frames), 384))
    ref = ref.permute(0, 2, 1).numpy()
    assert np.linalg.norm(raw - ref) / np.linalg.norm(ref) < 1e-2


def test_intermediate_features_close(small_engine):
    from oracle import prepost
    eng = small_engine
    frames = _frames(1, 512, 768, seed=4)
    eng.preprocess(frames)
    eng.detect(1, conf=0.25)
    m = _oracle_model(eng._sd)
    taps = {}
    with torch.no_grad():
        m(prepost.preprocess(list(frames), 384), taps)
    from geotrax_b200 import GtError
    for layer in (0, 1, 2, 4, 6, 9, 12, 15, 18, 21):
        try:
            got = eng.act_to_f32(eng.feature(layer, 1))
        except GtError:
            assert layer == 1, f"layer {layer} has no buffer"   # model.1 is chained into model.2.cv1: its output never exists in HBM
            continue
        ref = taps[str(layer)].permute(0, 2, 3, 1).numpy()
        rel = np.linalg.norm(got - ref) / np.linalg.norm(ref)
        print("layer", layer, "rel", rel)
        assert rel < 1e-2, f"layer {layer}: rel {rel}"


def _rand_pred(B, A, nc, rng, rotated=False, frac=0.05):
    xy = rng.uniform(0, 1900, (B, A, 2))
    wh = rng.uniform(10, 120, (B, A, 2))
    cls = rng.uniform(0, 0.2, (B, A, nc))
    hot = rng.random((B, A)) < frac
    cls[hot, rng.integers(0, nc, hot.sum())] = rng.uniform(0.25, 0.99, hot.sum())
    parts = [xy, wh, cls]
    if rotated:
        parts.append(rng.uniform(-0.7, 2.3, (B, A, 1)))
    return np.concatenate(parts, 2).astype(np.float32)


@pytest.mark.parametrize("agnostic", [True, False])
@pytest.mark.parametrize("A,frac", [(3000, 0.05), (3000, 0.0), (8064, 0.9)])
def test_nms_indices_bit_exact(mid_engine, agnostic, A, frac):
    """criterion (2): identical keep indices / classes; boxes equal (same decoded inputs on both sides)."""
    from oracle import prepost
    rng = np.random.default_rng(A + int(agnostic))
    pred = _rand_pred(2, A, 4, rng, frac=frac)
    # clustered boxes so that suppression actually happens
    pred[:, ::3, :2] = pred[:, 1::3, :2][:, : pred[:, ::3].shape[1]] + rng.uniform(-6, 6, pred[:, ::3, :2].shape).astype(np.float32)
    rows, counts, keep = mid_engine.nms(pred, 4, False, 0.25, 0.7, agnostic, [0, 1, 2, 3], 300)
    ref, idxs = prepost.non_max_suppression(torch.from_numpy(pred).permute(0, 2, 1), 0.25, 0.7, [0, 1, 2, 3], agnostic, 300, nc=4, return_idxs=True)
    for b in range(2):
        n = int(counts[b])
        assert n == len(ref[b]), f"image {b}: kept {n} vs oracle {len(ref[b])}"
        assert np.array_equal(keep[b, :n], idxs[b].numpy()), "keep indices differ"
        assert np.array_equal(rows[b, :n, 5], ref[b][:, 5].numpy())
        assert np.array_equal(rows[b, :n, :5], ref[b][:, :5].numpy()), "boxes/conf differ bitwise"


def test_nms_rotated_matches_oracle(mid_engine):
    small_engine = mid_engine
    from oracle import prepost
    rng = np.random.default_rng(7)
    pred = _rand_pred(2, 2500, 4, rng, rotated=True, frac=0.08)
    rows, counts, keep = small_engine.nms(pred, 4, True, 0.25, 0.5, True, None, 300)
    ref, idxs = prepost.non_max_suppression(torch.from_numpy(pred).permute(0, 2, 1), 0.25, 0.5, None, True, 300, nc=4, rotated=True, return_idxs=True)
    for b in range(2):
        n = int(counts[b])
        got, want = set(keep[b, :n].tolist()), set(idxs[b].numpy().tolist())
        # probiou uses transcendental functions: allow boundary flips on < 1 % of the kept set
        assert len(got ^ want) <= max(1, len(want) // 100), f"rotated keep sets differ: {len(got ^ want)} of {len(want)}"


def test_detect_end_to_end_matches_oracle_on_same_raw(small_engine):
    """GPU decode + filter + NMS + scale_boxes vs the oracle's, both fed the GPU's raw head tensor."""
    from oracle import prepost
    from oracle.yolov8 import YOLOv8
    eng = small_engine
    frames = _frames(2, 512, 768, seed=9)
    eng.preprocess(frames)
    boxes, counts, keep = eng.detect(2, conf=0.05, iou=0.7, agnostic=True, classes=[0, 1, 2, 3], want_keep=True)
    raw = torch.from_numpy(eng.raw_head(2)).permute(0, 2, 1).contiguous()
    head = YOLOv8(4).model[22]
    shapes = [(32, 48), (16, 24), (8, 12)]
    dec = head.decode(raw, shapes)
    outs, idxs = prepost.non_max_suppression(dec, 0.05, 0.7, [0, 1, 2, 3], True, 300, nc=4, return_idxs=True)
    assert counts.sum() > 10, "test needs some detections; lower conf"
    for b in range(2):
        n = int(counts[b])
        ref = outs[b].clone()
        ref[:, :4] = prepost.scale_boxes((256, 384), ref[:, :4], (512, 768))
        got_set, ref_set = set(keep[b, :n].tolist()), set(idxs[b].numpy().tolist())
        assert len(got_set ^ ref_set) <= max(1, len(ref_set) // 50), f"keep sets differ by {len(got_set ^ ref_set)} of {len(ref_set)}"
        common = [i for i in keep[b, :n] if i in ref_set]
        ref_by = {int(a): r for a, r in zip(idxs[b].numpy(), ref.numpy())}
        got_by = {int(a): r for a, r in zip(keep[b, :n], boxes[b, :n])}
        for a in common:
            assert got_by[a][5] == ref_by[a][5]
            np.testing.assert_allclose(got_by[a][:5], ref_by[a][:5], rtol=2e-4, atol=2e-3)


def test_chained_conv_equals_two_launches():
    """model.1 + model.2.cv1 as one chained kernel (GT_CHAIN=1, the default) against the two-launch form (GT_CHAIN=0): same 16-bit
    intermediate, same accumulation order -> the C2f output and the raw head must agree to the last bit."""
    import os
    from conftest import _make_engine
    frames = _frames(2, 512, 768, seed=9)
    outs = []
    for chain in ("1", "0"):
        old = os.environ.get("GT_CHAIN")
        os.environ["GT_CHAIN"] = chain
        try:
            eng = _make_engine("fp16")
        finally:
            if old is None:
                os.environ.pop("GT_CHAIN")
            else:
                os.environ["GT_CHAIN"] = old
        eng.preprocess(frames)
        eng.detect(2, conf=0.25)
        outs.append((eng.feature(2, 2).copy(), eng.raw_head(2).copy(), eng.conv_kernel_info()[0]))
        eng.close()
    (f1, r1, n1), (f0, r0, n0) = outs
    assert n1 == n0 - 1, f"chained graph should have one conv launch less ({n1} vs {n0})"
    assert np.array_equal(f1, f0), f"C2f output differs: {np.mean(f1 != f0):.2e} of the values"
    assert np.array_equal(r1, r0)
