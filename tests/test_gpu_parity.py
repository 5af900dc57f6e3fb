"""GPU parity tests (run on the B200 box: pytest -m gpu).  Everything goes through the C ABI (libbyolo.so); the
oracle is only the checker.  Tolerances are stated next to each comparison and explained in DESIGN.md."""
import os

import numpy as np
import pytest
import torch

import golden_inputs as GI
from golden_inputs import stress_rows
from byolo import priors as P
from byolo import weights as W
from oracle import decode as D
from oracle import net as ON
from oracle import nms as ONMS
from oracle import philox

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(__file__), 'golden')
PRI = P.as_scale_list(P.by_stride('ECP_9_PRIORS'))
OBJ = {'standard': 4, 'aleatoric': 9, 'epistemic': 14}


def _excess(a, b, rtol, atol):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    err = np.abs(a - b) - (atol + rtol * np.abs(b))
    err[np.isnan(a) & np.isnan(b)] = -1
    return err


def assert_close(a, b, rtol, atol, what):
    err = _excess(a, b, rtol, atol)
    bad = ~(err <= 0)
    assert not bad.any(), '%s: %d/%d out of tolerance, worst excess %g at %s (got %r want %r); bad per last-axis index: %r' % (
        what, bad.sum(), bad.size, np.nanmax(err), np.unravel_index(np.nanargmax(np.where(np.isnan(err), np.inf, err)), err.shape),
        np.asarray(a)[bad][:4], np.asarray(b)[bad][:4], dict(zip(*np.unique(np.nonzero(bad)[-1], return_counts=True))))


# ------------------------------------------------------------------------------------------------------ K3: NMS
def _nms_gpu(rows, obj_idx, max_out=1000, cluster=0, chunked=False):
    """cluster / chunked: byolo_nms_ex test hooks (CTAs per image, the N > 32768 kernel for any N)."""
    import byolo
    r = torch.from_numpy(np.ascontiguousarray(rows, np.float32)).cuda()
    boxes, cnt, idx = byolo.nms(r, obj_idx, max_out, cluster=cluster, chunked=chunked)
    torch.cuda.synchronize()
    return boxes.cpu().numpy(), cnt.cpu().numpy(), idx.cpu().numpy()


@pytest.mark.parametrize('name', list(GI.CASES))
def test_nms_bit_exact_on_reference_rows(name):
    """K3 fed the rows the reference code produced must return exactly what the reference code's NMS returned."""
    case, g = GI.CASES[name], np.load(os.path.join(G, name + '.npz'))
    boxes, cnt, idx = _nms_gpu(g['rows'], OBJ[case['variant']])
    assert np.array_equal(cnt, g['nms_count'])
    assert np.array_equal(boxes, g['nms_rows'])            # padded with zeros beyond count


@pytest.mark.parametrize('cluster', [0, 1, 2, 4, 8])
def test_nms_stress_full_size_bit_exact(cluster):
    """cluster = CTAs per image (0: the library's own choice); every split of the pair tests must select the same boxes."""
    rows = stress_rows(4, 105)
    boxes, cnt, idx = _nms_gpu(rows, 14, cluster=cluster)
    for b in range(rows.shape[0]):
        want = ONMS.nms(rows[b], 14)
        assert cnt[b] == len(want)
        assert np.array_equal(idx[b, :cnt[b]], want), 'image %d' % b
        assert np.array_equal(boxes[b, :cnt[b]], rows[b][want])
        assert np.all(idx[b, cnt[b]:] == -1) and np.all(boxes[b, cnt[b]:] == 0)
    assert cnt.min() == 1000                                    # generator guarantees >= 1000 survivors


def test_nms_packed_output_carries_the_count_row():
    """byolo_nms_ex(packed): [B, max_out + 1, D] with (count, 0, ...) in the extra row - the all-gather message."""
    import byolo
    rows = stress_rows(3, 107)
    rows[2, :, 14] = 0.5
    rows[2, :, :4] = [0.1, 0.1, 0.5, 0.5]                       # one survivor: the count row must say 1
    r = torch.from_numpy(rows).cuda()
    boxes, cnt, idx = byolo.nms(r, 14, 1000)
    packed, idx2 = byolo.nms(r, 14, 1000, packed=True)
    torch.cuda.synchronize()
    packed = packed.cpu().numpy()
    assert packed.shape == (3, 1001, 23)
    assert np.array_equal(packed[:, :1000], boxes.cpu().numpy()) and np.array_equal(idx2.cpu().numpy(), idx.cpu().numpy())
    assert np.array_equal(packed[:, 1000, 0].astype(np.int32), cnt.cpu().numpy()) and np.all(packed[:, 1000, 1:] == 0)
    assert packed[2, 1000, 0] == 1


def test_nms_per_class_matches_the_reference_variant():
    """inference_epistemic.py:104-126 (commented, "used to produce the results for the paper"): per class, rows whose class
    score is strictly greater than the other's -> NMS(1000) -> concatenated.  Bit exact vs the oracle NMS on the subsets."""
    import byolo
    rows = stress_rows(2, 108)
    rows[:, :, 17:19] = np.random.default_rng(9).random((2, rows.shape[1], 2), dtype=np.float32)     # class scores
    rows[0, :50, 18] = rows[0, :50, 17]                          # exact class ties belong to no class (tf.greater)
    rows[1, 2000:, 17] = 0.9                                     # image 1: few riders -> fewer than 1000 survive
    rows[1, 2000:, 18] = 0.1
    got = byolo.nms_per_class(torch.from_numpy(rows).cuda(), 14, 17, 2)
    for b in range(2):
        want = []
        for cls in (0, 1):
            sub = rows[b][rows[b][:, 17 + cls] > rows[b][:, 18 - cls]]
            want.append(sub[ONMS.nms(sub, 14)])
        want = np.concatenate(want)
        assert got[b].shape == want.shape and np.array_equal(got[b], want), (b, got[b].shape, want.shape)
    assert len(got[1]) < 2000


def _random_rows(B, N, D, obj_idx, seed, tie_frac=0.05):
    rng = np.random.default_rng(seed)
    rows = rng.random((B, N, D), dtype=np.float32)
    cy, cx = rng.random((B, N), dtype=np.float32), rng.random((B, N), dtype=np.float32)
    h, w = (0.02 + 0.1 * rng.random((B, N))).astype(np.float32), (0.01 + 0.05 * rng.random((B, N))).astype(np.float32)
    rows[..., 0], rows[..., 1], rows[..., 2], rows[..., 3] = cy - h / 2, cx - w / 2, cy + h / 2, cx + w / 2
    score = (1.0 / (1.0 + np.exp(-rng.normal(-3.0, 2.0, (B, N))))).astype(np.float32)
    n_tie = int(N * tie_frac)
    for b in range(B):                                           # exact score ties -> index order decides
        src, dst = rng.integers(0, N, n_tie), rng.integers(0, N, n_tie)
        score[b, dst] = score[b, src]
    rows[..., obj_idx] = score
    return rows


@pytest.mark.parametrize('cluster', [1, 8])
def test_nms_chunked_kernel_on_stress_rows(cluster):
    """The chunked kernel (score-ordered chunks of <= 4096 candidates, used for N > 32768) forced onto the 22743-row stress
    case: same selection as the oracle."""
    rows = stress_rows(2, 106)
    boxes, cnt, idx = _nms_gpu(rows, 14, cluster=cluster, chunked=True)
    for b in range(rows.shape[0]):
        want = ONMS.nms(rows[b], 14)
        assert cnt[b] == len(want) and np.array_equal(idx[b, :cnt[b]], want), 'image %d' % b
        assert np.array_equal(boxes[b, :cnt[b]], rows[b][want])


def test_nms_beyond_32768_candidates():
    """N = 120960 = the reference's ECP geometry (1024 x 1920, inference_epistemic.py:213-233), D = 23."""
    rows = _random_rows(2, 120960, 23, 14, 11)
    boxes, cnt, idx = _nms_gpu(rows, 14)
    for b in range(rows.shape[0]):
        want = ONMS.nms(rows[b], 14)
        assert cnt[b] == len(want) == 1000 and np.array_equal(idx[b, :cnt[b]], want), 'image %d' % b
        assert np.array_equal(boxes[b, :cnt[b]], rows[b][want])


def _tie_rows():
    """10000 candidates share ONE score and only 500 distinct boxes (the radix select has to descend into the index bits and
    three chunks are visited for ~500 survivors), then 6000 candidates with another single score and small distinct boxes."""
    rows = _random_rows(1, 16000, 7, 4, 12, tie_frac=0.0)
    rows[0, :, 2] = rows[0, :, 0] + 0.004
    rows[0, :, 3] = rows[0, :, 1] + 0.004
    rows[0, :10000, :4] = rows[0, np.arange(10000) % 500, :4]
    rows[0, :10000, 4] = 0.75
    rows[0, 10000:, 4] = 0.5
    return rows


def test_nms_chunked_with_more_ties_than_a_chunk():
    rows = _tie_rows()
    boxes, cnt, idx = _nms_gpu(rows, 4, max_out=2000, chunked=True)
    want = ONMS.nms(rows[0], 4, 2000)
    assert cnt[0] == len(want) == 2000 and np.array_equal(idx[0, :cnt[0]], want)
    assert 0 < (want < 10000).sum() <= 500 and (want >= 10000).sum() >= 1500


@pytest.mark.parametrize('cluster', [1, 8])
def test_nms_edge_cases(cluster):
    rng = np.random.default_rng(5)
    # all boxes identical -> 1 survivor; zero-area boxes never suppress; tiny N; max_out smaller than survivors
    rows = np.zeros((3, 70, 7), np.float32)
    rows[0, :, :4] = [0.1, 0.1, 0.5, 0.5]
    rows[0, :, 4] = rng.random(70)
    rows[1, :, :4] = [0.3, 0.3, 0.3, 0.9]                      # zero area
    rows[1, :, 4] = 0.5                                        # all tied -> index order
    rows[2, :, :2] = rng.random((70, 2))
    rows[2, :, 2:4] = rows[2, :, :2] + 0.01
    rows[2, :, 4] = rng.random(70)
    boxes, cnt, idx = _nms_gpu(rows, 4, max_out=50, cluster=cluster)
    for b in range(3):
        want = ONMS.nms(rows[b], 4, 50)
        assert cnt[b] == len(want) and np.array_equal(idx[b, :cnt[b]], want)
    assert cnt[0] == 1 and cnt[1] == 50 and np.array_equal(idx[1, :50], np.arange(50))


# ------------------------------------------------------------------------------------------------------ K2: decode
@pytest.mark.parametrize('name', list(GI.CASES))
def test_decode_on_reference_raw_outputs(name):
    import byolo
    case, g = GI.CASES[name], np.load(os.path.join(G, name + '.npz'))
    v = case['variant']
    H, Wd = case['img_size'][:2]
    eng = byolo.Engine(v, (H, Wd), case['cls_cnt'], T=case.get('T', 1), max_batch=case['batch'], precision='fp32')
    raws = []
    for j in range(3):
        r = g['raw%d' % j]
        raws.append(torch.from_numpy(np.ascontiguousarray(r.reshape((-1,) + r.shape[-3:]))).cuda())
    rows = eng.decode(raws, case['batch']).cpu().numpy()
    # fp32 decode vs the reference graph's fp32 decode: 1e-5 relative; cancellation columns get an absolute floor
    # (covariance diag 4:8 ~ eps * |t|^2, determinant 12 ~ round-off of a rank-deficient 4x4, mutual information 15/19)
    atol = np.full(rows.shape[-1], 2e-6)
    rtol = np.full(rows.shape[-1], 2e-5)
    if v == 'epistemic':
        atol[4:8] = 3e-5
        atol[[15, 19]] = 3e-6
        rtol[12], atol[12] = 2e-2, 1e-6        # det of the 4x4 covariance: LU of values that carry ~1e-6 cancellation noise
    assert_close(rows, g['rows'], rtol, atol, 'rows')


# ------------------------------------------------------------------------------------------------------ K1: conv layers
def ref_conv(x, kernel, bn=None, bias=None, x2=None, residual=None, stride=1, upsample=False, drop=None, emulate=False,
             t1=1, t2=1):
    """Oracle for byolo_conv_layer: torch CPU fp32 conv with the engine's folding; emulate=True rounds the operands
    (and the stored fp16 output) exactly where the fp16 paths round them.  t1 / t2: stack_feature_map of x / x2
    (tf.concat([x]*T, axis=0) per image, layers.py:595-597: sample s reads sample s // t)."""
    rd = (lambda t: t.half().float()) if emulate else (lambda t: t)
    if t1 > 1:
        x = np.repeat(x, t1, axis=0)
    if t2 > 1:
        x2 = np.repeat(x2, t2, axis=0)
    xin = torch.from_numpy(x if x2 is None else np.concatenate([x, x2], -1)).permute(0, 3, 1, 2)
    k = torch.from_numpy(kernel).permute(3, 2, 0, 1)
    if bn is not None:
        scale = torch.from_numpy(bn['gamma'] / np.sqrt(bn['var'] + np.float32(1e-5)))
        shift = torch.from_numpy(bn['beta']) - torch.from_numpy(bn['mean']) * scale
        k = k * scale.view(-1, 1, 1, 1)
    else:
        shift = torch.from_numpy(bias)
    xin, k = rd(xin), rd(k)
    if stride == 2:
        y = torch.nn.functional.conv2d(torch.nn.functional.pad(xin, (1, 1, 1, 1)), k, stride=2)
    else:
        y = torch.nn.functional.conv2d(xin, k, padding=kernel.shape[0] // 2)
    if drop is not None:
        seed, lid, T, image0, p = drop
        S, C, Ho, Wo = y.shape
        m = np.stack([philox.keep_mask(seed, lid, s % T, image0 + s // T, (Ho, Wo, C), p) for s in range(S)])
        y = y * (1.0 / (1.0 - p)) * torch.from_numpy(m).permute(0, 3, 1, 2).float()
    y = y + shift.view(1, -1, 1, 1)
    if bn is not None:
        y = torch.maximum(y, 0.1 * y)
        if residual is not None:
            y = y + rd(torch.from_numpy(residual).permute(0, 3, 1, 2))
        y = rd(y)
    if upsample:
        y = torch.nn.functional.interpolate(y, scale_factor=2, mode='nearest')
    return y.permute(0, 2, 3, 1).contiguous().numpy()


CONV_CASES = {
    # name: (S, H, W, c1, c2, k, stride, cout, residual, upsample, dropout, dense)
    'stem_3_32': (2, 32, 64, 3, 0, 3, 1, 32, False, False, False, False),
    'pw_64_32': (3, 20, 12, 64, 0, 1, 1, 32, False, False, False, False),
    'c3_32_64_sw64': (2, 16, 16, 32, 0, 3, 1, 64, False, False, False, False),
    'c3_64_128_res': (2, 19, 19, 64, 0, 3, 1, 128, True, False, False, False),
    'c3_128_256': (1, 38, 38, 128, 0, 3, 1, 256, False, False, False, False),
    'pw_1024_512_drop': (4, 8, 8, 1024, 0, 1, 1, 512, False, False, True, False),
    'c3_512_1024_drop': (2, 6, 6, 512, 0, 3, 1, 1024, False, False, True, False),
    's2_64_128': (3, 32, 32, 64, 0, 3, 2, 128, False, False, False, False),
    's2_32_64_sw64': (2, 24, 40, 32, 0, 3, 2, 64, False, False, False, False),
    's2_256_512_odd': (5, 12, 20, 256, 0, 3, 2, 512, False, False, False, False),
    'cat_128_256_drop': (4, 12, 20, 128, 256, 1, 1, 128, False, False, True, False),
    'pw_up': (2, 6, 10, 256, 0, 1, 1, 128, False, True, False, False),
    'det_42': (4, 12, 20, 256, 0, 1, 1, 42, False, False, False, True),
    'det_21': (2, 5, 3, 1024, 0, 1, 1, 21, False, False, False, True),
}
# MC-stacked sources (A_STACK1 / A_STACK2): name -> (conv case fields..., t1, t2); S is the number of samples the conv sees
STACK_CASES = {
    'stack1_1024_512_drop': ((6, 7, 5, 1024, 0, 1, 1, 512, False, False, True, False), 3, 1),     # conv "75": in1 = L74 of S/3 images
    'stack1_256_128': ((4, 9, 11, 256, 0, 1, 1, 128, False, False, False, False), 2, 1),
    'stack2_cat_256_512_drop': ((6, 10, 6, 256, 512, 1, 1, 256, False, False, True, False), 1, 3),  # conv "87": [upsampled, L61 stacked]
    'stack2_cat_128_256': ((8, 12, 20, 128, 256, 1, 1, 128, False, False, False, False), 1, 4),     # conv "99"
}


def _conv_case(name):
    t1 = t2 = 1
    if name in STACK_CASES:
        (S, H, Wd, c1, c2, k, stride, cout, res, up, drop, dense), t1, t2 = STACK_CASES[name]
    else:
        S, H, Wd, c1, c2, k, stride, cout, res, up, drop, dense = CONV_CASES[name]
    import zlib
    rng = np.random.default_rng(zlib.crc32(name.encode()))
    x = rng.standard_normal((S // t1, H, Wd, c1)).astype(np.float32)
    if c1 == 3:
        x = rng.random((S, H, Wd, 3), dtype=np.float32)          # the stem reads images in [0,1) (dataset_utils.py:6-11)
    x2 = rng.standard_normal((S // t2, H, Wd, c2)).astype(np.float32) if c2 else None
    cin = c1 + c2
    kernel = (rng.standard_normal((k, k, cin, cout)) * np.sqrt(2.0 / (k * k * cin))).astype(np.float32)
    bn = bias = None
    if dense:
        bias = rng.standard_normal(cout).astype(np.float32)
    else:
        bn = dict(beta=(rng.standard_normal(cout) * .1).astype(np.float32), gamma=rng.uniform(.8, 1.2, cout).astype(np.float32),
                  mean=(rng.standard_normal(cout) * .1).astype(np.float32), var=rng.uniform(.5, 1.5, cout).astype(np.float32))
    residual = rng.standard_normal((S, H // stride, Wd // stride, cout)).astype(np.float32) if res else None
    dropspec = (77, 3, max(t1, t2, 2), 5, 0.1) if drop else None          # seed, layer id, T, image0, p
    return dict(x=x, x2=x2, kernel=kernel, bn=bn, bias=bias, residual=residual, stride=stride, upsample=up, drop=dropspec,
                t1=t1, t2=t2)


def _run_conv(c, precision):
    import byolo
    t = lambda a: torch.from_numpy(a).cuda() if a is not None else None
    d = c['drop']
    out = byolo.conv_layer(t(c['x']), c['kernel'], bn=c['bn'], bias=c['bias'], x2=t(c['x2']), residual=t(c['residual']),
                           stride=c['stride'], upsample=c['upsample'], precision=precision,
                           dropout_layer=d[1] if d else -1, T=d[2] if d else 1, seed=d[0] if d else 0,
                           image_index0=d[3] if d else 0, drop_prob=d[4] if d else 0.1, t1=c['t1'], t2=c['t2'])
    torch.cuda.synchronize()
    return out.cpu().numpy()


def _ref(c, emulate=False):
    return ref_conv(c['x'], c['kernel'], c['bn'], c['bias'], c['x2'], c['residual'], c['stride'], c['upsample'], c['drop'],
                    emulate=emulate, t1=c['t1'], t2=c['t2'])


@pytest.mark.parametrize('name', list(CONV_CASES) + list(STACK_CASES))
def test_conv_fp32_cuda_core_path(name):
    """Exact path: fp32 operands and activations; tolerance 1e-4 (summation order only)."""
    c = _conv_case(name)
    got = _run_conv(c, 'fp32')
    want = _ref(c)
    assert got.shape == want.shape
    assert_close(got, want, 1e-4, 1e-4, name)


@pytest.mark.parametrize('name', list(CONV_CASES) + list(STACK_CASES))
def test_conv_split_fp16_tensor_core_path(name):
    """tcgen05 split-fp16 mode (hi*hi + hi*lo + lo*hi, fp32 accumulate, activations stored as hi + lo pairs) vs the
    plain fp32 oracle: the same tolerance as the fp32 CUDA-core path (summation order only)."""
    c = _conv_case(name)
    got = _run_conv(c, 'fp16x3')
    want = _ref(c)
    assert got.shape == want.shape
    assert_close(got, want, 1e-4, 1e-4, name)


@pytest.mark.parametrize('name', list(CONV_CASES) + list(STACK_CASES))
def test_conv_fp16_tensor_core_path(name):
    """tcgen05 path vs the oracle with operands rounded to fp16 where the kernel rounds them.  fp16 outputs may differ
    by one fp16 ulp (2^-10 relative) where the fp32 sums straddle a rounding boundary: rtol 2e-3, atol 2e-3."""
    c = _conv_case(name)
    got = _run_conv(c, 'fp16')
    want = _ref(c, emulate=True)
    assert got.shape == want.shape
    assert_close(got, want, 2e-3, 2e-3, name)
    simt = _run_conv(c, 'fp16-simt')                        # CUDA-core twin with identical rounding points
    assert_close(got, simt, 2e-3, 2e-3, name + ' vs fp16-simt')


# ------------------------------------------------------------------------------------------------------ end to end
def _engine_for(case, precision):
    import byolo
    H, Wd = case['img_size'][:2]
    eng = byolo.Engine(case['variant'], (H, Wd), case['cls_cnt'], T=case.get('T', 1), max_batch=case['batch'],
                       precision=precision)
    eng.load_weights(W.synthetic(case['variant'], case['cls_cnt'], case['weight_seed']))
    return eng


def _oracle_rows(case, emulate=None, keep=False):
    w = W.synthetic(case['variant'], case['cls_cnt'], case['weight_seed'])
    fwd = ON.Forward(case['variant'], w, case['cls_cnt'], torch.float32, emulate=emulate, keep_layers=keep)
    res = fwd.run(GI.images(case), T=case.get('T'), seed=case.get('dropout_seed', 0))
    if case['variant'] == 'epistemic':
        rows = np.stack([D.rows_from_raw('epistemic', r['raw'], PRI) for r in res])
    else:
        rows = D.rows_from_raw(case['variant'], res[0]['raw'], PRI)
    return rows, res


def _first_bad_layer(eng, res, case, rtol, atol):
    """Walks the 75 conv outputs and reports the first one that deviates (diagnostic for failures)."""
    msgs = []
    for i in range(75):
        got = eng.activation(i).cpu().numpy()
        want = np.concatenate([r['conv_out'][i] for r in res])      # image-major, then MC sample: s = b*T + t
        if got.shape[1] == 2 * want.shape[1]:        # convs 84/96 store through the fused nearest x2 upsample
            want = want.repeat(2, axis=1).repeat(2, axis=2)
        if got.shape != want.shape:
            return 'conv %d: shape %s vs %s' % (i, got.shape, want.shape)
        err = _excess(got, want, rtol, atol)
        if (err > 0).any():
            msgs.append('conv %d: %d/%d bad, max abs diff %.3g' % (i, (err > 0).sum(), err.size, np.abs(got - want).max()))
            if len(msgs) >= 3:
                break
    return '; '.join(msgs) if msgs else 'all conv outputs within tolerance'


@pytest.mark.parametrize('precision', ['fp32', 'fp16x3'])
@pytest.mark.parametrize('name', list(GI.CASES))
def test_forward_matches_reference_graph(name, precision):
    """Whole path in the exact precision modes - fp32 CUDA cores, and the split-fp16 tensor-core mode - vs the golden rows
    produced by the reference's own graph code.  Tolerance: the north-star 1e-3 relative (observed ~1e-5), absolute
    floors for cancellation columns."""
    case, g = GI.CASES[name], np.load(os.path.join(G, name + '.npz'))
    eng = _engine_for(case, precision)
    img = torch.from_numpy(GI.images(case)).cuda()
    boxes, cnt, idx, rows = eng.detect(img, seed=case.get('dropout_seed', 0), want_rows=True)
    torch.cuda.synchronize()
    rows = rows.cpu().numpy()
    atol = np.full(rows.shape[-1], 1e-4)
    rtol = np.full(rows.shape[-1], 1e-3)
    if case['variant'] == 'epistemic':
        atol[4:8] = 1e-3
        rtol[12], atol[12] = 5e-2, 1e-5
    err = _excess(rows, g['rows'], rtol, atol)
    if case['variant'] == 'epistemic':
        # det(cov) inherits the ~1e-6 absolute cancellation noise of the covariance entries times the cofactors:
        # bounded by a fraction of prod(diag) (det <= prod(diag) for a PSD matrix); the reference calls the
        # column "not useful" (inference_epistemic.py:157)
        err[..., 12] = np.minimum(err[..., 12], np.abs(rows[..., 12] - g['rows'][..., 12]) - 0.05 * np.prod(np.abs(g['rows'][..., 4:8]), -1))
    if (err > 0).any():
        _, res = _oracle_rows(case, keep=True)
        pytest.fail('rows out of tolerance (%d/%d, per column %r); %s' % (
            (err > 0).sum(), err.size, dict(zip(*np.unique(np.nonzero(err > 0)[-1], return_counts=True))),
            _first_bad_layer(eng, res, case, 1e-3, 1e-3)))
    # NMS on the engine's own rows == oracle NMS on the same rows (bit exact), and close to the reference's selection
    cnt, idx, boxes = cnt.cpu().numpy(), idx.cpu().numpy(), boxes.cpu().numpy()
    agree = []
    for b in range(case['batch']):
        want = ONMS.nms(rows[b], OBJ[case['variant']])
        assert cnt[b] == len(want) and np.array_equal(idx[b, :cnt[b]], want)
        assert np.array_equal(boxes[b, :cnt[b]], rows[b][want])
        ref_sel = ONMS.nms(g['rows'][b], OBJ[case['variant']])
        n = min(len(want), len(ref_sel))
        agree.append(np.mean(want[:n] == ref_sel[:n]))
    assert min(agree) > 0.98, agree


@pytest.mark.parametrize('name', list(GI.CASES))
def test_forward_fp16_tensor_core_path(name):
    """Product path (tcgen05, fp16 operands).  (1) against the oracle run with the same operand rounding: tight;
    (2) against the fp32 reference rows: the error of fp16 operands themselves, bounded statistically."""
    case, g = GI.CASES[name], np.load(os.path.join(G, name + '.npz'))
    eng = _engine_for(case, 'fp16')
    img = torch.from_numpy(GI.images(case)).cuda()
    rows = eng.forward(img, seed=case.get('dropout_seed', 0))
    torch.cuda.synchronize()
    rows = rows.cpu().numpy()
    want, res = _oracle_rows(case, emulate='half', keep=True)
    # (1a) every conv output vs the oracle with identical rounding points.  Two valid fp16 computations differ where
    # an fp32 sum straddles an fp16 rounding boundary (1 ulp = 2^-10 relative) and such flips propagate and compound
    # with depth, so single layers are pinned by the conv-layer tests above; here: no gross outlier in any layer and
    # the deviation stays at the few-ulp level (median) through all 75 layers.
    stats = []
    for i in range(75):
        got = eng.activation(i).cpu().numpy()
        w = np.concatenate([r['conv_out'][i] for r in res])
        if got.shape[1] == 2 * w.shape[1]:
            w = w.repeat(2, axis=1).repeat(2, axis=2)
        assert got.shape == w.shape, (i, got.shape, w.shape)
        scale = np.abs(w).max() + 1e-6
        d = np.abs(got - w)
        stats.append((i, float(np.median(d / (np.abs(w) + 1e-2 * scale))), float((_excess(got, w, 4e-3, 4e-3) > 0).mean()),
                      float(d.max() / scale)))
    if os.environ.get('BYOLO_DIAG_DIR'):
        with open(os.path.join(os.environ['BYOLO_DIAG_DIR'], 'fp16_layer_stats_%s.txt' % name), 'w') as f:
            f.write('conv  median_rel  frac>4e-3  max_abs/scale\n')
            for st in stats:
                f.write('%3d  %.3e  %.4f  %.3e\n' % st)
    for i, med, frac, mx in stats:
        assert mx < 0.05, 'conv %d: gross outlier, max |diff| = %.3g of the layer scale' % (i, mx)
        assert med < 2e-3, 'conv %d: median relative deviation %.3g' % (i, med)
    # (1b) final rows vs the same oracle, (2) vs the fp32 reference graph: median relative error (cancellation columns
    # excluded) must stay at the fp16 operand-rounding level measured in DESIGN.md (~4e-4 .. 8e-4).
    cols = [c for c in range(rows.shape[-1]) if not (case['variant'] == 'epistemic' and c in (4, 5, 6, 7, 12, 15, 19))]
    rel_emu = np.abs(rows - want) / (np.abs(want) + 1e-2)
    rel_ref = np.abs(rows - g['rows']) / (np.abs(g['rows']) + 1e-2)
    assert np.nanmedian(rel_emu[..., cols]) < 1e-3, np.nanmedian(rel_emu[..., cols])
    assert np.nanmedian(rel_ref[..., cols]) < 2e-3, np.nanmedian(rel_ref[..., cols])
    assert np.nanquantile(rel_ref[..., cols], 0.99) < 5e-2, np.nanquantile(rel_ref[..., cols], 0.99)
