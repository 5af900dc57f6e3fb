"""Size-independent properties at the BASELINE.json geometries (608x608 / 416x416), where the CPU oracle is too slow to
run per test: determinism, shard invariance of the dropout stream, degenerate dropout, agreement between the exact and
the tensor-core path, NMS structure."""
import numpy as np
import pytest
import torch

from byolo import weights as W

pytestmark = pytest.mark.gpu


def _engine(variant, size, T=1, B=2, precision='fp16', drop_prob=0.1, seed=0):
    import byolo
    eng = byolo.Engine(variant, (size, size), 2, T=T, max_batch=B, precision=precision, drop_prob=drop_prob)
    return eng.load_weights(W.synthetic(variant, 2, seed))


def _images(B, size, seed):
    return torch.from_numpy(np.random.default_rng(seed).random((B, size, size, 3), dtype=np.float32)).cuda()


def test_epistemic_608_deterministic_and_seed_dependent():
    eng = _engine('epistemic', 608, T=4, B=2)
    img = _images(2, 608, 103)
    a = eng.forward(img, seed=7).clone()
    b = eng.forward(img, seed=7).clone()
    c = eng.forward(img, seed=8).clone()
    torch.cuda.synchronize()
    assert a.shape == (2, 22743, 23)
    assert torch.equal(torch.nan_to_num(a), torch.nan_to_num(b))               # same masks in, same numbers out (bitwise)
    assert not torch.equal(torch.nan_to_num(a[..., :4]), torch.nan_to_num(c[..., :4]))
    assert torch.all(a[..., 21] >= 0) and torch.all(a[..., 21] <= 2) and torch.all(a[..., 22] <= 2)   # layer / prior ids


def test_shard_invariance_of_the_dropout_stream_608():
    """Images processed as one batch or as two shards with image_index0 offsets give identical rows: what makes the
    gathered multi-GPU result equal the single-GPU result."""
    eng = _engine('epistemic', 608, T=3, B=4)
    img = _images(4, 608, 104)
    whole = eng.forward(img, seed=11).clone()
    lo = eng.forward(img[:2].contiguous(), seed=11, image_index0=0).clone()
    hi = eng.forward(img[2:].contiguous(), seed=11, image_index0=2).clone()
    torch.cuda.synchronize()
    assert torch.equal(torch.nan_to_num(whole[:2]), torch.nan_to_num(lo))
    assert torch.equal(torch.nan_to_num(whole[2:]), torch.nan_to_num(hi))
    wrong = eng.forward(img[2:].contiguous(), seed=11, image_index0=0)
    assert not torch.equal(torch.nan_to_num(whole[2:, :, :4]), torch.nan_to_num(wrong[:, :, :4]))


def test_zero_dropout_collapses_to_the_aleatoric_path_416():
    """drop_prob = 0: every MC sample is identical, so the epistemic covariance is exactly 0, mutual information 0 and
    boxes / scores / aleatoric variances equal the aleatoric model's (same weights)."""
    w = W.synthetic('epistemic', 2, 5)
    import byolo
    epi = byolo.Engine('epistemic', (416, 416), 2, T=5, max_batch=1, precision='fp32', drop_prob=0.0).load_weights(w)
    ale = byolo.Engine('aleatoric', (416, 416), 2, max_batch=1, precision='fp32').load_weights(w)
    img = _images(1, 416, 105)
    e = epi.forward(img).cpu().numpy()[0]
    a = ale.forward(img).cpu().numpy()[0]
    assert e.shape == (10647, 23) and a.shape == (10647, 16)
    assert np.allclose(e[:, :4], a[:, :4], rtol=2e-5, atol=2e-6)               # boxes (mean of T equal values: <= 1 ulp off)
    assert np.allclose(e[:, 8:12], a[:, 4:8], rtol=2e-5)                       # aleatoric variances (mean of T equal values)
    assert np.allclose(e[:, 14], a[:, 9], rtol=2e-5, atol=1e-7) and np.allclose(e[:, 17:19], a[:, 11:13], rtol=2e-5, atol=1e-7)
    assert np.abs(e[:, 4:8]).max() < 1e-3 * (1 + np.abs(e[:, :4]).max())       # E[xx] - E[x]^2 of identical samples: round-off only
    mi = e[:, [15, 19]]
    assert np.nanmax(np.abs(mi)) < 1e-5


@pytest.mark.parametrize('variant,size', [('aleatoric', 608), ('epistemic', 416)])
def test_tensor_core_path_tracks_the_exact_path_full_size(variant, size):
    """Config 2 / config 4 geometry: fp16 tcgen05 path vs fp32 CUDA-core path on the same inputs (both on the GPU)."""
    T = 3 if variant == 'epistemic' else 1
    img = _images(1, size, 106)
    rows = {}
    for prec in ('fp32', 'fp16'):
        rows[prec] = _engine(variant, size, T=T, B=1, precision=prec, seed=9).forward(img, seed=3).cpu().numpy()[0]
    cols = [c for c in range(rows['fp32'].shape[1]) if not (variant == 'epistemic' and c in (4, 5, 6, 7, 12, 15, 19))]
    rel = np.abs(rows['fp16'] - rows['fp32']) / (np.abs(rows['fp32']) + 1e-2)
    assert np.nanmedian(rel[:, cols]) < 2e-3, np.nanmedian(rel[:, cols])
    assert np.nanquantile(rel[:, cols], 0.99) < 5e-2


def test_detect_output_structure_608():
    """NMS output invariants: scores non-increasing in selection order, no kept pair above the IoU threshold, padding."""
    eng = _engine('epistemic', 608, T=2, B=2)
    boxes, cnt, idx, rows = eng.detect(_images(2, 608, 107), seed=1, want_rows=True)
    torch.cuda.synchronize()
    boxes, cnt, idx, rows = boxes.cpu().numpy(), cnt.cpu().numpy(), idx.cpu().numpy(), rows.cpu().numpy()
    for b in range(2):
        n = cnt[b]
        assert 0 < n <= 1000 and np.all(idx[b, n:] == -1) and np.all(boxes[b, n:] == 0)
        assert np.array_equal(boxes[b, :n], rows[b][idx[b, :n]], equal_nan=True)          # entropies are NaN at saturated scores, like the reference
        s = boxes[b, :n, 14]
        assert np.all(s[:-1] >= s[1:])
        bx = boxes[b, :min(n, 300), :4].astype(np.float64)
        y0, x0, y1, x1 = (np.minimum(bx[:, 0], bx[:, 2]), np.minimum(bx[:, 1], bx[:, 3]), np.maximum(bx[:, 0], bx[:, 2]),
                          np.maximum(bx[:, 1], bx[:, 3]))
        area = (y1 - y0) * (x1 - x0)
        ih = np.maximum(np.minimum(y1[:, None], y1[None]) - np.maximum(y0[:, None], y0[None]), 0)
        iw = np.maximum(np.minimum(x1[:, None], x1[None]) - np.maximum(x0[:, None], x0[None]), 0)
        inter = ih * iw
        iou = inter / (area[:, None] + area[None] - inter + 1e-30)
        np.fill_diagonal(iou, 0)
        assert iou.max() <= 0.5 + 1e-6


def test_detect_on_a_wide_frame_with_more_than_32768_anchors():
    """608 x 1216 (non-square, N = 45486 > 32768): the whole path runs and the chunked NMS picks exactly what the oracle
    picks from the engine's own rows.  (The reference's ECP geometry, 1024 x 1920, has N = 120960: covered for the NMS
    alone in test_gpu_parity.py::test_nms_beyond_32768_candidates.)"""
    import byolo
    from oracle import nms as ONMS
    eng = byolo.Engine('epistemic', (608, 1216), 2, T=2, max_batch=1, precision='fp16').load_weights(W.synthetic('epistemic', 2, 0))
    img = torch.from_numpy(np.random.default_rng(5).random((1, 608, 1216, 3), dtype=np.float32)).cuda()
    boxes, cnt, idx, rows = eng.detect(img, seed=3, want_rows=True)
    torch.cuda.synchronize()
    rows, cnt, idx = rows.cpu().numpy(), cnt.cpu().numpy(), idx.cpu().numpy()
    assert rows.shape == (1, 45486, 23)
    want = ONMS.nms(rows[0], 14)
    assert cnt[0] == len(want) and np.array_equal(idx[0, :cnt[0]], want)


def test_reference_main_configuration_ecp_1024x1920_T50():
    """The configuration inference_epistemic.main() ships with (inference_epistemic.py:212-233): one 1024 x 1920 frame,
    T = 50 MC samples, cls_cnt 2 -> 120960 anchors through the chunked NMS.  Checks structure, determinism and that the
    selection equals the oracle's on the engine's own rows; prints the time per frame."""
    import byolo
    from oracle import nms as ONMS
    eng = byolo.Engine('epistemic', (1024, 1920), 2, T=50, max_batch=1, precision='fp16').load_weights(W.synthetic('epistemic', 2, 0))
    img = torch.from_numpy(np.random.default_rng(9).random((1, 1024, 1920, 3), dtype=np.float32)).cuda()
    boxes, cnt, idx, rows = eng.detect(img, seed=5, want_rows=True)
    boxes2, cnt2, idx2 = eng.detect(img, seed=5)
    torch.cuda.synchronize()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(3):
        eng.detect(img, seed=5)
    t1.record()
    torch.cuda.synchronize()
    print('ECP frame 1024x1920, T=50: %.2f ms per frame' % (t0.elapsed_time(t1) / 3))
    rows, cnt_h, idx_h = rows.cpu().numpy(), cnt.cpu().numpy(), idx.cpu().numpy()
    assert rows.shape == (1, 120960, 23) and boxes.shape == (1, 1000, 23)
    assert torch.equal(idx, idx2) and torch.equal(torch.nan_to_num(boxes), torch.nan_to_num(boxes2))
    want = ONMS.nms(rows[0], 14)
    assert cnt_h[0] == len(want) and np.array_equal(idx_h[0, :cnt_h[0]], want)
