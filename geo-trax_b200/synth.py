"""Seed-deterministic synthetic bird's-eye-view frames (SURVEY.md section 8d "synthetic frame generator").

The reference's sample clip ``data/U_video_cut.mp4`` is absent from the mount (/root/reference/.MISSING_LARGE_BLOBS) and
there is no network, so parity tests and bench.py use frames generated here: a multi-scale textured ground plane with
road markings, warped per frame by a known small homography (the golden drift envelope: a few px of translation,
<= 0.1 deg rotation, perspective <= 1e-6 -- BASELINE.md section 1), plus 100-200 filled rectangles ("vehicles", golden box
statistics: w 60-110 px, h 30-50 px) that move linearly and therefore do NOT follow the background homography.
"""
from __future__ import annotations

from typing import List, Tuple

import cv2
import numpy as np


def _texture(h: int, w: int, rng: np.random.Generator) -> np.ndarray:
    acc = np.zeros((h, w), np.float32)
    amp = 1.0
    for cell in (256, 64, 16, 4):
        gh, gw = max(2, h // cell + 2), max(2, w // cell + 2)
        g = rng.standard_normal((gh, gw)).astype(np.float32)
        acc += amp * cv2.resize(g, (w, h), interpolation=cv2.INTER_CUBIC)
        amp *= 0.6
    acc += 0.35 * rng.standard_normal((h, w)).astype(np.float32)
    acc = (acc - acc.mean()) / (acc.std() + 1e-6)
    return acc


def make_ground(h: int, w: int, seed: int = 0) -> np.ndarray:
    """u8 BGR ground plane of size (h, w)."""
    rng = np.random.default_rng(seed)
    base = _texture(h, w, rng)
    img = np.empty((h, w, 3), np.float32)
    for c, (m, s) in enumerate(((105, 34), (110, 36), (112, 38))):
        img[..., c] = m + s * base + 6.0 * cv2.GaussianBlur(rng.standard_normal((h, w)).astype(np.float32), (0, 0), 1.5)
    img = np.clip(img, 0, 255).astype(np.uint8)
    # road markings: long light lines + dashed segments
    n_lines = max(4, w // 320)
    for i in range(n_lines):
        y = int(rng.integers(h // 10, h - h // 10))
        cv2.line(img, (0, y), (w - 1, y + int(rng.integers(-h // 20, h // 20 + 1))), (215, 215, 215), max(1, h // 540))
        x = int(rng.integers(w // 10, w - w // 10))
        for y0 in range(0, h, max(8, h // 27)):
            cv2.line(img, (x, y0), (x, y0 + max(4, h // 54)), (225, 225, 225), max(1, h // 540))
    return img


def make_boxes(n: int, h: int, w: int, rng: np.random.Generator, scale: float = 1.0) -> np.ndarray:
    """(n, 4) xywh f32, golden-like vehicle sizes at 4K (scaled for smaller frames)."""
    bw = rng.uniform(60, 110, n) * scale
    bh = rng.uniform(30, 50, n) * scale
    flip = rng.random(n) < 0.35
    bw, bh = np.where(flip, bh, bw), np.where(flip, bw, bh)
    xc = rng.uniform(bw / 2 + 2, w - bw / 2 - 2, n)
    yc = rng.uniform(bh / 2 + 2, h - bh / 2 - 2, n)
    return np.stack([xc, yc, bw, bh], 1).astype(np.float32)


def draw_vehicles(img: np.ndarray, boxes: np.ndarray, rng: np.random.Generator) -> None:
    for xc, yc, bw, bh in boxes:
        col = tuple(int(v) for v in rng.integers(20, 250, 3))
        p0 = (int(round(xc - bw / 2)), int(round(yc - bh / 2)))
        p1 = (int(round(xc + bw / 2)), int(round(yc + bh / 2)))
        cv2.rectangle(img, p0, p1, col, -1)
        cv2.rectangle(img, (p0[0] + 3, p0[1] + 3), (p1[0] - 3, p1[1] - 3), tuple(min(255, c + 35) for c in col), 1)


def small_homography(rng: np.random.Generator, h: int, w: int, max_t: float = 6.0, max_rot_deg: float = 0.1, max_persp: float = 1e-6):
    """current -> reference homography inside the golden drift envelope (full-resolution pixel units)."""
    t = rng.uniform(-max_t, max_t, 2)
    a = np.deg2rad(rng.uniform(-max_rot_deg, max_rot_deg))
    s = 1.0 + rng.uniform(-5e-4, 5e-4)
    cx, cy = w / 2.0, h / 2.0
    R = np.array([[s * np.cos(a), -s * np.sin(a), 0], [s * np.sin(a), s * np.cos(a), 0], [0, 0, 1.0]])
    C = np.array([[1, 0, cx], [0, 1, cy], [0, 0, 1.0]])
    Ci = np.array([[1, 0, -cx], [0, 1, -cy], [0, 0, 1.0]])
    H = C @ R @ Ci
    H[0, 2] += t[0]
    H[1, 2] += t[1]
    H[2, 0], H[2, 1] = rng.uniform(-max_persp, max_persp, 2)
    return H / H[2, 2]


def make_flight(n_frames: int, h: int = 2160, w: int = 3840, seed: int = 0, n_vehicles: int = 132, noise: float = 2.0,
                max_t: float = 6.0) -> Tuple[List[np.ndarray], List[np.ndarray], List[np.ndarray]]:
    """-> (frames u8 BGR, boxes xywh per frame, H_gt per frame (current -> reference; identity for frame 0))."""
    rng = np.random.default_rng(seed)
    ground = make_ground(h, w, seed)
    scale = w / 3840.0
    boxes0 = make_boxes(n_vehicles, h, w, rng, scale)
    vel = rng.uniform(-3.0, 3.0, (n_vehicles, 2)).astype(np.float32) * scale
    frames, boxes, Hs = [], [], []
    for t in range(n_frames):
        H = np.eye(3) if t == 0 else small_homography(rng, h, w, max_t * scale)
        # p_ref = H p_cur  =>  cur(p) = ground(H p): warpPerspective with WARP_INVERSE_MAP uses dst(p) = src(M p)
        img = ground.copy() if t == 0 else cv2.warpPerspective(ground, H, (w, h), flags=cv2.INTER_LINEAR | cv2.WARP_INVERSE_MAP,
                                                                borderMode=cv2.BORDER_REFLECT_101)
        b = boxes0.copy()
        b[:, :2] += vel * t
        b[:, 0] = np.clip(b[:, 0], b[:, 2] / 2 + 1, w - b[:, 2] / 2 - 1)
        b[:, 1] = np.clip(b[:, 1], b[:, 3] / 2 + 1, h - b[:, 3] / 2 - 1)
        draw_vehicles(img, b, np.random.default_rng(seed + 7))  # same colours every frame
        if noise > 0:
            nz = rng.standard_normal((h, w, 1)).astype(np.float32) * noise
            img = np.clip(img.astype(np.float32) + nz, 0, 255).astype(np.uint8)
        frames.append(img)
        boxes.append(b)
        Hs.append(H)
    return frames, boxes, Hs


def make_frames(n: int, h: int = 2160, w: int = 3840, seed: int = 0, n_vehicles: int = 132) -> np.ndarray:
    """(n, h, w, 3) u8 BGR batch (bench input)."""
    frames, _, _ = make_flight(n, h, w, seed, n_vehicles)
    return np.stack(frames)


def bgr_to_nv12(frame_bgr: np.ndarray) -> np.ndarray:
    """BGR24 -> NV12 (u8 [H * 3 / 2][W]: Y plane, then interleaved U, V rows) through cv2: synthetic input for the decoder-format ingest."""
    import cv2
    h, w = frame_bgr.shape[:2]
    i420 = cv2.cvtColor(frame_bgr, cv2.COLOR_BGR2YUV_I420)
    flat, q = i420.reshape(-1), (h // 2) * (w // 2)
    out = np.empty((h * 3 // 2, w), np.uint8)
    out[:h] = i420[:h]
    out[h:] = np.stack([flat[h * w:h * w + q].reshape(h // 2, w // 2), flat[h * w + q:h * w + 2 * q].reshape(h // 2, w // 2)], -1).reshape(h // 2, w)
    return out


# ---- a tiny H.264 elementary-stream writer (lossless I_PCM macroblocks): input for the NVDEC ingest path ---------------------------------
class _Bits:
    def __init__(self):
        self.bits = []

    def u(self, n, v):
        self.bits.extend((v >> (n - 1 - i)) & 1 for i in range(n))

    def ue(self, v):
        v += 1
        n = v.bit_length()
        self.u(n - 1, 0)
        self.u(n, v)

    def se(self, v):
        self.ue(2 * v - 1 if v > 0 else -2 * v)

    def trailing(self):
        self.bits.append(1)
        while len(self.bits) % 8:
            self.bits.append(0)

    def align_zero(self):
        while len(self.bits) % 8:
            self.bits.append(0)

    def bytes(self):
        assert len(self.bits) % 8 == 0
        return bytes(int("".join(map(str, self.bits[i:i + 8])), 2) for i in range(0, len(self.bits), 8))


def _escape(payload: bytes) -> bytes:
    """Emulation prevention (H.264 7.4.1): a 0x03 after every two zero bytes that are followed by a byte <= 3 (small payloads only)."""
    out, zeros = bytearray(), 0
    for b in payload:
        if zeros >= 2 and b <= 3:
            out.append(3)
            zeros = 0
        out.append(b)
        zeros = zeros + 1 if b == 0 else 0
    return bytes(out)


def _nal(nal_ref_idc: int, nal_type: int, payload: bytes) -> bytes:
    return b"\x00\x00\x00\x01" + bytes([(nal_ref_idc << 5) | nal_type]) + payload


def h264_ipcm_stream(nv12_frames, level_idc: int = 52) -> Tuple[bytes, np.ndarray]:
    """Annex-B H.264 elementary stream (Constrained Baseline, CAVLC, every picture an IDR of I_PCM macroblocks, deblocking off) that
    decodes EXACTLY to the given NV12 frames -- a lossless bitstream any H.264 decoder (NVDEC included) accepts, written without an
    encoder library (there is none in this image, and no network).  Frame dimensions must be multiples of 16.

    H.264 forbids the PCM sample value 0 (it could emulate a start code), so samples are clamped to [1, 255]; the returned array is
    the clamped NV12 the decoder must reproduce bit for bit.  -> (stream bytes, expected NV12 u8 [n][H * 3 / 2][W])."""
    frames = np.maximum(np.asarray(nv12_frames, np.uint8), 1)
    n, h32, w = frames.shape
    h = h32 * 2 // 3
    assert h % 16 == 0 and w % 16 == 0, "I_PCM writer: frame dimensions must be multiples of 16"
    mbw, mbh = w // 16, h // 16
    sps = _Bits()
    sps.u(8, 66); sps.u(8, 0xC0); sps.u(8, level_idc)        # profile_idc 66 (baseline), constraint_set0/1 flags, level
    sps.ue(0)                                                 # seq_parameter_set_id
    sps.ue(0)                                                 # log2_max_frame_num_minus4
    sps.ue(2)                                                 # pic_order_cnt_type 2: output order = decode order
    sps.ue(1); sps.u(1, 0)                                    # max_num_ref_frames, gaps_in_frame_num_value_allowed_flag
    sps.ue(mbw - 1); sps.ue(mbh - 1)
    sps.u(1, 1); sps.u(1, 1); sps.u(1, 0); sps.u(1, 1)        # frame_mbs_only, direct_8x8_inference, frame_cropping, vui_parameters_present
    # VUI: only the bitstream restriction -- no reordering, one frame of DPB -- so that decoders output every picture immediately
    sps.u(1, 0); sps.u(1, 0); sps.u(1, 0); sps.u(1, 0); sps.u(1, 0)   # aspect_ratio, overscan, video_signal_type, chroma_loc, timing info: absent
    sps.u(1, 0); sps.u(1, 0); sps.u(1, 0)                     # nal_hrd, vcl_hrd parameters absent, pic_struct_present_flag
    sps.u(1, 1)                                               # bitstream_restriction_flag
    sps.u(1, 1); sps.ue(0); sps.ue(0); sps.ue(16); sps.ue(16) # motion_vectors_over_pic_boundaries, max_bytes_per_pic_denom, max_bits_per_mb_denom, log2_max_mv_length_h / v
    sps.ue(0); sps.ue(1)                                      # max_num_reorder_frames, max_dec_frame_buffering
    sps.trailing()
    pps = _Bits()
    pps.ue(0); pps.ue(0); pps.u(1, 0); pps.u(1, 0)            # pps id, sps id, entropy_coding_mode (CAVLC), bottom_field_pic_order_in_frame_present
    pps.ue(0); pps.ue(0); pps.ue(0)                           # num_slice_groups_minus1, num_ref_idx_l0/l1_default_active_minus1
    pps.u(1, 0); pps.u(2, 0)                                  # weighted_pred_flag, weighted_bipred_idc
    pps.se(0); pps.se(0); pps.se(0)                           # pic_init_qp_minus26, pic_init_qs_minus26, chroma_qp_index_offset
    pps.u(1, 1); pps.u(1, 0); pps.u(1, 0)                     # deblocking_filter_control_present, constrained_intra_pred, redundant_pic_cnt_present
    pps.trailing()
    out = [_nal(3, 7, _escape(sps.bytes())), _nal(3, 8, _escape(pps.bytes()))]
    for i in range(n):
        y = frames[i, :h].reshape(mbh, 16, mbw, 16).transpose(0, 2, 1, 3).reshape(mbh * mbw, 256)
        uv = frames[i, h:].reshape(h // 2, w // 2, 2)
        cb = uv[..., 0].reshape(mbh, 8, mbw, 8).transpose(0, 2, 1, 3).reshape(mbh * mbw, 64)
        cr = uv[..., 1].reshape(mbh, 8, mbw, 8).transpose(0, 2, 1, 3).reshape(mbh * mbw, 64)
        hdr = _Bits()
        hdr.ue(0); hdr.ue(7); hdr.ue(0)                       # first_mb_in_slice, slice_type 7 (I, all slices of the picture), pic_parameter_set_id
        hdr.u(4, 0)                                           # frame_num (IDR)
        hdr.ue(i & 1)                                         # idr_pic_id: differs between consecutive IDR pictures
        hdr.u(1, 0); hdr.u(1, 0)                              # no_output_of_prior_pics_flag, long_term_reference_flag
        hdr.se(0)                                             # slice_qp_delta
        hdr.ue(1)                                             # disable_deblocking_filter_idc = 1
        hdr.ue(25)                                            # mb_type I_PCM of macroblock 0
        hdr.align_zero()                                      # pcm_alignment_zero_bit
        mb = np.empty((mbh * mbw, 386), np.uint8)
        mb[:, 0], mb[:, 1] = 0x0D, 0x00                       # ue(25) = 000011010 + 7 alignment zero bits (the stream is byte aligned after PCM samples)
        mb[:, 2:258], mb[:, 258:322], mb[:, 322:386] = y, cb, cr
        # macroblock 0's mb_type lives in the slice header bits; samples are >= 1 and each later macroblock header is 0x0D 0x00, so the
        # sample part never holds two consecutive zero bytes: only the few header bytes need the emulation-prevention pass
        body = mb.reshape(-1)[2:]
        assert not ((body[:-1] == 0) & (body[1:] == 0)).any()
        out.append(_nal(3, 5, _escape(hdr.bytes()) + body.tobytes() + b"\x80"))   # + rbsp trailing bits
    return b"".join(out), frames
