"""ORACLE (test infrastructure, never shipped): Philox4x32-10 dropout-mask stream.

The reference's dropout is `tf.layers.dropout(rate=0.1, training=True)` with no seed
(/root/reference/lib_yolo/layers.py:521-524), i.e. not reproducible.  Parity for the MC-dropout path is
therefore defined as "same masks in, same numbers out": the CUDA epilogue and this file implement the SAME
counter-based stream, specified here.

Stream specification (shared with bayesian-yolov3_b200/csrc/philox.cuh):
  * generator: Philox4x32-10 (Salmon et al., SC'11), multipliers 0xD2511F53 / 0xCD9E8D57,
    Weyl key increments 0x9E3779B9 / 0xBB67AE85.
  * key     = (seed & 0xffffffff, seed >> 32)
  * counter = (group, layer_id, t, image)   with  group = element_index // 8,
    element_index = (y*W + x)*C + c  over the UN-padded [H,W,C] map of one sample,
    layer_id = 0..14 in execution order of the dropout-bearing head convs,
    t = MC sample index, image = GLOBAL image index (so sharding over ranks does not change masks).
  * one call yields 4x32 bits = 8 x 16-bit lanes; element (group*8 + j) uses word j//2, low half for even j,
    high half for odd j.
  * keep  iff  r16 >= thr16,  thr16 = round(drop_prob * 65536)   (drop_prob=0.1 -> 6554/65536 = 0.1000061)
  * kept values are scaled by 1/(1-drop_prob) in fp32 (TF: x * (1/keep_prob) * mask, Appendix B-4 of SURVEY.md).
"""
import numpy as np

M0 = np.uint64(0xD2511F53)
M1 = np.uint64(0xCD9E8D57)
W0 = 0x9E3779B9
W1 = 0xBB67AE85
MASK32 = np.uint64(0xFFFFFFFF)


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """Vectorised Philox4x32-10. Counters are uint32 arrays (broadcastable); key words are python ints.
    Returns four uint32 arrays."""
    c0 = np.asarray(c0, dtype=np.uint64)
    c1 = np.asarray(c1, dtype=np.uint64)
    c2 = np.asarray(c2, dtype=np.uint64)
    c3 = np.asarray(c3, dtype=np.uint64)
    c0, c1, c2, c3 = np.broadcast_arrays(c0, c1, c2, c3)
    k0 = int(k0) & 0xFFFFFFFF
    k1 = int(k1) & 0xFFFFFFFF
    for _ in range(10):
        p0 = M0 * c0            # 32x32 -> 64 bit products (exact in uint64)
        p1 = M1 * c2
        hi0, lo0 = p0 >> np.uint64(32), p0 & MASK32
        hi1, lo1 = p1 >> np.uint64(32), p1 & MASK32
        n0 = hi1 ^ c1 ^ np.uint64(k0)
        n1 = lo1
        n2 = hi0 ^ c3 ^ np.uint64(k1)
        n3 = lo0
        c0, c1, c2, c3 = n0, n1, n2, n3
        k0 = (k0 + W0) & 0xFFFFFFFF
        k1 = (k1 + W1) & 0xFFFFFFFF
    return (c0.astype(np.uint32), c1.astype(np.uint32), c2.astype(np.uint32), c3.astype(np.uint32))


def drop_threshold16(drop_prob):
    return int(round(float(drop_prob) * 65536.0))


def rand16(seed, layer_id, t, image, n_elements):
    """The first n_elements 16-bit lanes of the (layer_id, t, image) stream."""
    assert n_elements % 8 == 0
    groups = np.arange(n_elements // 8, dtype=np.uint64)
    w = philox4x32_10(groups, layer_id, t, image, seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    out = np.empty((n_elements // 8, 8), dtype=np.uint32)
    for i in range(4):
        out[:, 2 * i] = w[i] & np.uint32(0xFFFF)
        out[:, 2 * i + 1] = w[i] >> np.uint32(16)
    return out.reshape(-1)


def keep_mask(seed, layer_id, t, image, shape_hwc, drop_prob):
    """Boolean keep mask of shape [H,W,C] for one sample of one dropout layer."""
    h, w, c = shape_hwc
    r = rand16(seed, layer_id, t, image, h * w * c)
    return (r >= drop_threshold16(drop_prob)).reshape(h, w, c)
