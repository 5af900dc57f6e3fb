"""Per-column parity statistics of detection rows against the fp32 oracle, and NMS selection agreement.

Used by tests/test_gpu_fullsize.py (BASELINE.json geometries) and tools/parity_table.py.  The tolerance model is the
north-star one: |got - want| <= atol[c] + rtol * |want| per element, rtol = 1e-3, with a per-column absolute floor
`atol[c]` that is written down here once:

  * coordinates, variances, scores, entropies: 1e-4 (values are O(1e-2..1), so the floor only matters near zero);
  * epistemic covariance diagonal (E[x^2] - E[x]^2 in fp32: cancellation noise ~ eps * |t|^2): 1e-3;
  * det of the 4x4 covariance: bounded relative to prod(diag) (det <= prod(diag) for a PSD matrix; the reference calls
    the column "not useful", /root/reference/inference_epistemic.py:157);
  * mutual information (difference of two entropies): 1e-4.
"""
import numpy as np

COLUMNS = {
    'standard': ['y0', 'x0', 'y1', 'x1', 'obj', 'cls0', 'cls1'],
    'aleatoric': ['y0', 'x0', 'y1', 'x1', 'var_x', 'var_y', 'var_w', 'var_h', 'prod_var', 'obj', 'obj_H', 'cls0', 'cls1', 'cls_H',
                  'layer', 'prior'],
    'epistemic': ['y0', 'x0', 'y1', 'x1', 'epi_x', 'epi_y', 'epi_w', 'epi_h', 'ale_x', 'ale_y', 'ale_w', 'ale_h', 'det_cov',
                  'sum_ale', 'obj', 'obj_MI', 'obj_H', 'cls0', 'cls1', 'cls_MI', 'cls_H', 'layer', 'prior'],
}
OBJ = {'standard': 4, 'aleatoric': 9, 'epistemic': 14}
RTOL = 1e-3


def floors(variant):
    a = np.full(len(COLUMNS[variant]), 1e-4)
    if variant == 'epistemic':
        a[4:8] = 1e-3
    return a


NAN_PARTNER_MAX = 1e-3


def excess(got, want, variant, rtol=RTOL):
    """> 0 where an element is outside atol[c] + rtol*|want|.  Entropy / mutual-information columns are NaN in the reference
    when a probability saturates to exactly 0 or 1 (p*log p, /root/reference/lib_yolo/layers.py:349-358): NaN == NaN is
    equal, and a NaN on ONE side only (the other side's probability is one ulp short of saturation) is accepted iff the
    finite partner is < 1e-3 in magnitude, i.e. the entropy term it stands for is at its limit value 0."""
    got, want = np.asarray(got, np.float64), np.asarray(want, np.float64)
    err = np.abs(got - want) - (floors(variant) + rtol * np.abs(want))
    if variant == 'epistemic':      # determinant column: see module docstring
        bound = 0.05 * np.prod(np.abs(want[..., 4:8]), -1) + 1e-12
        err[..., 12] = np.minimum(err[..., 12], np.abs(got[..., 12] - want[..., 12]) - bound)
    gn, wn = np.isnan(got), np.isnan(want)
    err[gn & wn] = -1
    one = gn ^ wn
    partner = np.where(gn, want, got)
    err[one] = np.where(np.abs(partner[one]) < NAN_PARTNER_MAX, -1, np.inf)
    return err


def column_table(got, want, variant):
    """list of dict(col, name, median, p99, max, frac_out, nan_both, nan_one): relative error |d| / (|want| + floor) per
    column over the elements that are finite on both sides."""
    got, want = np.asarray(got, np.float64), np.asarray(want, np.float64)
    fl = floors(variant)
    ex = excess(got, want, variant)
    out = []
    for c, name in enumerate(COLUMNS[variant]):
        g, w = got[..., c].ravel(), want[..., c].ravel()
        ok = ~(np.isnan(g) | np.isnan(w))
        rel = np.abs(g[ok] - w[ok]) / (np.abs(w[ok]) + fl[c])
        out.append(dict(col=c, name=name, median=float(np.median(rel)), p99=float(np.quantile(rel, 0.99)), max=float(rel.max()),
                        max_abs=float(np.abs(g[ok] - w[ok]).max()), frac_out=float((ex[..., c] > 0).mean()),
                        nan_both=int((np.isnan(g) & np.isnan(w)).sum()), nan_one=int((np.isnan(g) ^ np.isnan(w)).sum())))
    return out


def nms_agreement(sel_got, sel_want):
    """Selection agreement of two NMS index lists (selection order): overlap = |A & B| / |A | B|, same_prefix = rank of
    the first divergence (== len if identical), same_rank = fraction of positions holding the same index."""
    a, b = list(map(int, sel_got)), list(map(int, sel_want))
    n = min(len(a), len(b))
    first = next((i for i in range(n) if a[i] != b[i]), n)
    sa, sb = set(a), set(b)
    return dict(n_got=len(a), n_want=len(b), overlap=len(sa & sb) / max(len(sa | sb), 1), first_divergence=first,
                same_rank=float(np.mean([a[i] == b[i] for i in range(n)])) if n else 1.0)


def format_table(title, table, agreements=None):
    lines = [title, '%-3s %-9s %10s %10s %10s %10s %10s %9s' % ('col', 'name', 'median', 'p99', 'max', 'max_abs', 'frac>tol', 'NaN b/1')]
    for r in table:
        lines.append('%-3d %-9s %10.2e %10.2e %10.2e %10.2e %10.2e %5d/%-3d' % (r['col'], r['name'], r['median'], r['p99'], r['max'],
                                                                              r['max_abs'], r['frac_out'], r['nan_both'], r['nan_one']))
    for i, a in enumerate(agreements or []):
        lines.append('image %d NMS: kept %d vs %d, index overlap %.4f, first divergence at rank %d, same index at same rank %.4f' % (
            i, a['n_got'], a['n_want'], a['overlap'], a['first_divergence'], a['same_rank']))
    return '\n'.join(lines) + '\n'
