"""CPU restatement of the arithmetic the tensor-core matchers rely on (geo-trax_b200/csrc/match_tc.cu) -- no GPU, numpy only:

* Hamming: key = popc(t)*8192 + j - 16384*<q, t> evaluated in float32 is exact, and the two smallest keys + popc(q)*8192 decode to
  cv2.BFMatcher's (distance, index) including ties and the 0 / 256 extremes;
* the operand image desc_expand_kernel writes is the canonical K-major SWIZZLE_128B tile (byte-address bits [4:6] ^= bits [7:9]),
  every byte of a group written exactly once;
* L2: ranking by |t|^2 - 2<q, t> on fp16-rounded operands keeps the true two nearest inside the best four for RootSIFT-like data."""
import cv2
import numpy as np


def _bits(d):
    return np.unpackbits(d, axis=1, bitorder="little").astype(np.float32)     # element k = bit k of the descriptor (byte k / 8, bit k % 8)


def test_hamming_keys_are_exact_in_float32_and_decode_to_bfmatcher():
    rng = np.random.default_rng(1)
    train = rng.integers(0, 256, (700, 32), dtype=np.uint8)
    query = train[rng.integers(0, 700, 300)] ^ (rng.integers(0, 256, (300, 32), dtype=np.uint8) & rng.integers(0, 256, (300, 32), dtype=np.uint8))
    train[100] = train[50]; train[200] = train[50]; train[7] = 0; train[8] = 255
    query[0] = 0; query[1] = 255; query[2] = ~train[5]
    Q, T = _bits(query), _bits(train)
    dot = (Q @ T.T).astype(np.float32)                                         # integers <= 256: exact
    c = (T.sum(1) * np.float32(8192) + np.arange(len(T), dtype=np.float32)).astype(np.float32)
    key = (np.float32(-16384) * dot + c[None, :]).astype(np.float32)           # every intermediate is an integer below 2^24
    assert np.array_equal(key, key.astype(np.int64).astype(np.float32)) and np.abs(key).max() < 2 ** 22
    order = np.argsort(key, axis=1, kind="stable")[:, :2]
    pa = (Q.sum(1) * np.float32(8192)).astype(np.float32)
    k12 = (np.take_along_axis(key, order, 1) + pa[:, None]).astype(np.int64)
    dist, idx = k12 >> 13, k12 & 8191
    ref = cv2.BFMatcher(cv2.NORM_HAMMING).knnMatch(query, train, k=2)
    ri = np.array([[m.trainIdx, n.trainIdx] for m, n in ref]); rd = np.array([[int(m.distance), int(n.distance)] for m, n in ref])
    assert np.array_equal(dist, rd) and np.array_equal(idx, ri)
    assert dist[0, 0] == 0 and Q.sum(1)[2] + T.sum(1)[5] - 2 * dot[2, 5] == 256      # the two extremes of the distance range occur


def test_operand_image_is_the_canonical_128_byte_swizzle():
    """desc_expand_kernel: byte (row r, k-block kb, 16-byte chunk cc, byte b) of a 128-row group sits at kb * 16384 + r * 128 + ((cc ^ (r & 7)) << 4) + b."""
    KB = 2
    seen = np.zeros(KB * 16384, np.int32)
    for r in range(128):
        for c in range(8 * KB):                           # 16 chunks of 16 bytes per row for E4M3 (256 bytes)
            kb, cc = c >> 3, c & 7
            off = kb * 16384 + r * 128 + ((cc ^ (r & 7)) << 4)
            logical = kb * 16384 + r * 128 + cc * 16                             # un-swizzled K-major address
            canonical = logical ^ (((logical >> 7) & 7) << 4)                    # Swizzle<3, 4, 3>: bits [4:6] ^= bits [7:9]
            assert off == canonical
            seen[off:off + 16] += 1
    assert (seen == 1).all()


def test_fp16_candidate_pass_keeps_the_true_two_nearest_in_its_top_four():
    rng = np.random.default_rng(2)

    def rootsift_like(n):
        d = rng.gamma(0.6, 18.0, (n, 128)).astype(np.float32)
        d[rng.random((n, 128)) < 0.35] = 0
        d = np.minimum(np.floor(d), 255).astype(np.float32); d[:, 0] += 1
        d /= d.sum(1, keepdims=True) + 1e-8
        return np.sqrt(d)

    train, query = rootsift_like(3000), rootsift_like(800)
    src = rng.choice(3000, 400, replace=False)
    noisy = np.maximum(train[src] + rng.normal(0, 0.01, (400, 128)).astype(np.float32), 0)
    query[:400] = noisy / np.linalg.norm(noisy, axis=1, keepdims=True)
    q16, t16 = query.astype(np.float16).astype(np.float32), train.astype(np.float16).astype(np.float32)
    key = (train ** 2).sum(1)[None, :] - 2.0 * (q16 @ t16.T)                     # what the tensor-core epilogue ranks by
    top4 = np.argsort(key, axis=1, kind="stable")[:, :4]
    d2 = ((query[:, None, :] - train[None, :, :]) ** 2).sum(-1)
    true2 = np.argsort(d2, axis=1, kind="stable")[:, :2]
    inside = [(set(true2[i]) <= set(top4[i])) for i in range(len(query))]
    assert all(inside), f"{len(inside) - sum(inside)} queries lose a true neighbour in the fp16 candidate pass"
    assert (true2[:400, 0] == src).all()
