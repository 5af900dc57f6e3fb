"""Pins the oracle (oracle/net.py, decode.py, nms.py) against outputs of the REFERENCE'S OWN code
(tests/golden/*.npz, produced by tests/golden/gen_golden.py from /root/reference on the numpy TF stand-in)."""
import os

import numpy as np
import pytest
import torch

import golden_inputs as GI
from byolo import priors as P
from byolo import weights as W
from oracle import decode as D
from oracle import net as ON
from oracle import nms as ONMS

G = os.path.join(os.path.dirname(__file__), 'golden')
PRI = P.as_scale_list(P.by_stride('ECP_9_PRIORS'))
OBJ = {'standard': 4, 'aleatoric': 9, 'epistemic': 14}


def _close(a, b, rtol, atol, what):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    err = np.abs(a - b) - (atol + rtol * np.abs(b))
    assert np.all((err <= 0) | (np.isnan(a) & np.isnan(b))), '%s: worst excess %g at %s' % (
        what, np.nanmax(err), np.unravel_index(np.nanargmax(err), err.shape))


@pytest.fixture(scope='module')
def runs():
    out = {}
    for name, case in GI.CASES.items():
        w = W.synthetic(case['variant'], case['cls_cnt'], case['weight_seed'])
        fwd = ON.Forward(case['variant'], w, case['cls_cnt'], torch.float32, keep_layers=True)
        out[name] = fwd.run(GI.images(case), T=case.get('T'), seed=case.get('dropout_seed', 0))
    return out


def test_prior_tables_match_reference():
    g = np.load(os.path.join(G, 'priors.npz'))
    for n in P.NAMES:
        mine = np.array([p for s in P.as_scale_list(P.by_stride(n)) for p in s])
        assert np.array_equal(mine, g[n]), n


def test_decode_matches_reference_numpy_implementation():
    """lib_yolo/utils.py:72-123 is the one non-TF statement of the aleatoric decode in the reference."""
    g = np.load(os.path.join(G, 'utils_numpy_decode.npz'))
    pri = [tuple(p) for p in g['priors']]
    rows = D.decode_aleatoric(g['pred'], pri, 0)                       # [S, 3*g*g, 16] prior-major
    S, lh, lw, _ = g['pred'].shape
    mine = rows.reshape(S, 3, lh, lw, 16).transpose(0, 2, 3, 1, 4).reshape(S, lh * lw * 3, 16)   # cell-major like utils
    ref = g['boxes']      # [y0,x0,y1,x1, var*4, obj, obj_stddev, cls*2, cls_stddev*2]
    _close(mine[..., 0:8], ref[..., 0:8], 2e-6, 1e-7, 'box+var')
    _close(mine[..., 9], ref[..., 8], 2e-6, 1e-7, 'obj')
    _close(mine[..., 11:13], ref[..., 10:12], 2e-6, 1e-7, 'cls')


@pytest.mark.parametrize('name', list(GI.CASES))
def test_forward_matches_reference_graph(name, runs):
    case, g = GI.CASES[name], np.load(os.path.join(G, name + '.npz'))
    res = runs[name]
    last = res[-1]
    # backbone taps (independent conv implementations: torch/oneDNN here, numpy im2col+sgemm in the shim)
    _close(last['layers'][74][-g['dn_out'].shape[0]:], g['dn_out'], 1e-4, 1e-4, 'dn_out')
    _close(last['layers'][36][-g['l36'].shape[0]:, ::4, ::4, ::8], g['l36'], 1e-4, 1e-4, 'l36')
    _close(last['layers'][61][-g['l61'].shape[0]:, ::2, ::2, ::8], g['l61'], 1e-4, 1e-4, 'l61')
    for j in range(3):
        mine = np.stack([r['raw'][j] for r in res]) if case['variant'] == 'epistemic' else res[0]['raw'][j]
        _close(mine, g['raw%d' % j], 1e-4, 2e-4, 'raw%d' % j)


@pytest.mark.parametrize('name', list(GI.CASES))
def test_decode_rows_match_reference_graph(name):
    """Decode fed with the reference's raw head outputs: isolates decode.py + concat order."""
    case, g = GI.CASES[name], np.load(os.path.join(G, name + '.npz'))
    v = case['variant']
    if v == 'epistemic':
        rows = np.stack([D.rows_from_raw(v, [g['raw%d' % j][b] for j in range(3)], PRI) for b in range(case['batch'])])
        # cancellation columns (covariance diag 4:8, det 12, MI 15/19): absolute floors, see DESIGN.md
        atol = np.full(23, 1e-6)
        atol[4:8] = 2e-5
        atol[12] = 1e-6
        atol[[15, 19]] = 2e-6
        _close(rows, g['rows'], 1e-4, atol, 'rows')
        rows64 = np.stack([D.rows_from_raw(v, [g['raw%d' % j][b].astype(np.float64) for j in range(3)], PRI,
                                           dtype=np.float64) for b in range(case['batch'])])
        a64 = np.full(23, 1e-5)
        a64[12] = 1e-6
        _close(rows64, g['rows64'], 2e-3, a64, 'rows64')       # raw inputs are the fp32 run's; fp64 run differs slightly
    else:
        rows = D.rows_from_raw(v, [g['raw%d' % j] for j in range(3)], PRI)
        _close(rows, g['rows'], 1e-5, 1e-6, 'rows')


@pytest.mark.parametrize('name', list(GI.CASES))
def test_nms_matches_reference_graph_bit_exact(name):
    """NMS fed with the reference's rows: selection and gathered rows must be identical."""
    case, g = GI.CASES[name], np.load(os.path.join(G, name + '.npz'))
    for b in range(case['batch']):
        got, idx = ONMS.nms_gather(g['rows'][b], OBJ[case['variant']])
        n = int(g['nms_count'][b])
        assert len(idx) == n
        assert np.array_equal(got, g['nms_rows'][b, :n])
        assert np.array_equal(idx, ONMS.nms_numpy(g['rows'][b], OBJ[case['variant']]))


def test_nms_ties_and_degenerate_boxes():
    rng = np.random.default_rng(3)
    n = 400
    c = rng.uniform(0, 1, (n, 2))
    s = rng.uniform(0.02, 0.3, (n, 2))
    rows = np.concatenate([c - s / 2, c + s / 2, rng.uniform(0, 1, (n, 1))], 1).astype(np.float32)
    rows[50:100, 4] = rows[0:50, 4]            # exact score ties -> lower index first
    rows[10, 2:4] = rows[10, 0:2]              # zero-area box: IoU 0 with everything, always selected
    rows[11, [0, 2]] = rows[11, [2, 0]]        # flipped corners are re-ordered by min/max
    a, b = ONMS.nms(rows, 4, 1000), ONMS.nms_numpy(rows, 4, 1000)
    assert np.array_equal(a, b) and 10 in a
    assert len(ONMS.nms(rows[:0], 4)) == 0
    assert len(ONMS.nms(rows, 4, max_out=7)) == 7


# ------------------------------------------------------------------------------------------------ independent pins
def test_philox_known_answers():
    """Philox4x32-10 known-answer vectors of the Random123 distribution (kat_vectors: counter, key -> output)."""
    from oracle import philox
    kat = [((0x00000000,) * 4, (0x00000000, 0x00000000), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
           ((0xffffffff,) * 4, (0xffffffff, 0xffffffff), (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
           ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0), (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1))]
    for ctr, key, want in kat:
        got = philox.philox4x32_10(*[np.array([c], np.uint32) for c in ctr], key[0], key[1])
        assert tuple(int(g[0]) for g in got) == want, (ctr, key)


def test_dropout_stream_statistics():
    """keep rate 0.9 (threshold 6554/65536), independent across MC samples, layers and images."""
    from oracle import philox
    shape = (38, 38, 256)
    m = [philox.keep_mask(1003, 4, t, 0, shape, 0.1) for t in range(4)]
    n = m[0].size
    for x in m:
        assert abs(x.mean() - (1 - 6554 / 65536)) < 4 * np.sqrt(0.09 / n)
    for a in range(4):
        for b in range(a + 1, 4):                      # P(both kept) = 0.81 if independent
            assert abs((m[a] & m[b]).mean() - 0.81) < 5e-3
    other_layer = philox.keep_mask(1003, 5, 0, 0, shape, 0.1)
    other_image = philox.keep_mask(1003, 4, 0, 1, shape, 0.1)
    other_seed = philox.keep_mask(1004, 4, 0, 0, shape, 0.1)
    for x in (other_layer, other_image, other_seed):
        assert abs((m[0] & x).mean() - 0.81) < 5e-3 and not np.array_equal(m[0], x)
    assert np.array_equal(m[0], philox.keep_mask(1003, 4, 0, 0, shape, 0.1))          # and reproducible


def test_nms_oracle_agrees_with_torchvision_on_tie_free_inputs():
    """torchvision.ops.nms implements the same greedy rule (suppress iff IoU > thr); on inputs without score ties and
    without degenerate boxes the kept index sequences must coincide (SURVEY.md 8c-iii sanity check)."""
    import torch
    import torchvision
    rng = np.random.default_rng(21)
    n = 3000
    cy, cx = rng.random(n), rng.random(n)
    h, w = 0.02 + 0.1 * rng.random(n), 0.02 + 0.1 * rng.random(n)
    rows = np.zeros((n, 7), np.float32)
    rows[:, 0], rows[:, 1], rows[:, 2], rows[:, 3] = cy - h / 2, cx - w / 2, cy + h / 2, cx + w / 2
    rows[:, 4] = rng.permutation(n).astype(np.float32) / n                            # distinct scores
    mine = ONMS.nms(rows, 4, 1000)
    xyxy = torch.from_numpy(rows[:, [1, 0, 3, 2]].copy())
    theirs = torchvision.ops.nms(xyxy, torch.from_numpy(rows[:, 4].copy()), 0.5)[:1000].numpy()
    agree = np.mean(mine[:len(theirs)] == theirs[:len(mine)])
    assert len(mine) == len(theirs) and agree > 0.995, (len(mine), len(theirs), agree)   # fp rounding of IoU near 0.5 may flip a pair


def test_epistemic_determinant_column_against_float64():
    """Column 12 of the epistemic row is det(cov) (layers.py:488): the oracle's fp32 value against a float64 evaluation of
    the same raw samples, bounded relative to prod(diag) (the scale of a 4x4 PSD determinant)."""
    rng = np.random.default_rng(5)
    raw = (rng.standard_normal((6, 3, 4, 42)) * 0.7).astype(np.float32)              # T = 6 samples of one 3x4 map
    rows32 = D.decode_epistemic(raw, PRI[1], 1, 2, np.float32)
    rows64 = D.decode_epistemic(raw.astype(np.float64), PRI[1], 1, 2, np.float64)
    scale = np.prod(np.abs(rows64[:, 4:8]), -1)
    assert np.all(np.abs(rows32[:, 12] - rows64[:, 12]) <= 1e-3 * scale + 1e-12)
    assert np.allclose(rows32[:, 4:8], rows64[:, 4:8], rtol=1e-3, atol=1e-6)          # the diagonal itself


def test_split_fp16_operands_carry_fp32_grade_precision():
    """The arithmetic model of the tensor-core split mode (precision='fp16x3'): operands rounded to hi + lo fp16 pairs.
    On the golden cases the rows stay within 1e-3 (+ floors) of the reference graph in EVERY element, while plain fp16
    operands leave it - the reason the mode exists.  (The accumulation order of the device - chunked
    round-to-nearest sums - is measured on the GPU, tests/test_gpu_fullsize.py.)"""
    import parity_report as PR
    for name in ('aleatoric_128', 'epistemic_96x160'):
        case, g = GI.CASES[name], np.load(os.path.join(G, name + '.npz'))
        w = W.synthetic(case['variant'], case['cls_cnt'], case['weight_seed'])
        out = {}
        for mode in ('split', 'half'):
            res = ON.Forward(case['variant'], w, case['cls_cnt'], torch.float32, emulate=mode).run(
                GI.images(case), T=case.get('T'), seed=case.get('dropout_seed', 0))
            out[mode] = (np.stack([D.rows_from_raw('epistemic', r['raw'], PRI) for r in res]) if case['variant'] == 'epistemic'
                         else D.rows_from_raw(case['variant'], res[0]['raw'], PRI))
        assert not (PR.excess(out['split'], g['rows'], case['variant']) > 0).any(), name
        assert (PR.excess(out['half'], g['rows'], case['variant']) > 0).mean() > 0.01, name
    # the pair itself: 22 significant bits while lo is a normal fp16 number (|x| >= 2^-3), an absolute error <= 2^-25 below
    # that (lo subnormal: the unbiased noise floor measured in profiles/r02/x3_probe_accumulator_truncation.txt)
    split_err = lambda v: (v.half().float() + (v - v.half().float()).half().float() - v).abs()
    big = (1 + torch.rand(1 << 16)) * torch.exp2(torch.randint(-3, 10, (1 << 16,)).float())
    assert float((split_err(big) / big).max()) <= 2.0 ** -21
    small = torch.rand(1 << 16) * 2.0 ** -3
    assert float(split_err(small).max()) <= 2.0 ** -25
