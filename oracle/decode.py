"""ORACLE (test infrastructure, never shipped): numpy restatement of the head-output split and the three decodes.

Follows /root/reference/lib_yolo/layers.py:
  split_detection            :11-38      split_detection_aleatoric  :41-84
  decode_bbox_standard       :191-258    decode_bbox_aleatoric      :261-346
  logistic_entropy / softmax_entropy :349-358
  decode_epistemic           :361-411    decode_bbox_epistemic      :414-502
and concat_bbox of /root/reference/inference_standard_yolov3.py:172-183, inference_aleatoric.py:181-192,
inference_epistemic.py:173-184.

All arithmetic is done in `dtype` (float32 reproduces the reference graph's precision; float64 is the
high-precision variant used to put a floor under cancellation-prone columns).  Prior tables are passed in as
[(h, w)] * 3 per scale (yolov3.py:29-61).
"""
import numpy as np


def sigmoid(x):
    return 1 / (1 + np.exp(-x))


def softmax(x):
    e = np.exp(x - np.max(x, axis=-1, keepdims=True))      # Appendix B-6: max-subtracted
    return e / np.sum(e, axis=-1, keepdims=True)


def logistic_entropy(s):                                   # layers.py:349-353 (NaN at s in {0,1} is reference behaviour)
    with np.errstate(divide='ignore', invalid='ignore'):
        no_obj = (1 - s) * np.log(1 - s)
        obj = s * np.log(s)
    return -(no_obj + obj)


def softmax_entropy(s):                                    # layers.py:356-358
    with np.errstate(divide='ignore', invalid='ignore'):
        return -np.sum(s * np.log(s), axis=-1)


def split(raw, cls_cnt, aleatoric):
    """raw [S,g,g,3*block] -> dict of loc[S,g,g,3,4], obj[S,g,g,3], cls[S,g,g,3,C] (+ log-variances)."""
    S, lh, lw, ch = raw.shape
    block = (2 * (5 + cls_cnt)) if aleatoric else (5 + cls_cnt)
    assert ch == 3 * block
    r = raw.reshape(S, lh, lw, 3, block)
    if not aleatoric:
        return {'loc': r[..., 0:4], 'obj': r[..., 4], 'cls': r[..., 5:5 + cls_cnt]}
    return {'loc': r[..., 0:4], 'log_loc_var': r[..., 4:8], 'obj': r[..., 8], 'log_obj_stddev': r[..., 9],
            'cls': r[..., 10:10 + cls_cnt], 'log_cls_stddev': r[..., 10 + cls_cnt:10 + 2 * cls_cnt]}


def _boxes(loc, priors, lh, lw, dtype):
    """loc [...,g,g,3,4] (t-space) -> y0,x0,y1,x1 each [...,g,g,3]."""
    x_off = np.arange(lw, dtype=dtype).reshape(1, lw, 1)
    y_off = np.arange(lh, dtype=dtype).reshape(lh, 1, 1)
    pw = np.array([p[1] for p in priors], dtype=dtype)
    ph = np.array([p[0] for p in priors], dtype=dtype)
    x = (x_off + sigmoid(loc[..., 0])) / dtype(lw)
    y = (y_off + sigmoid(loc[..., 1])) / dtype(lh)
    w = np.exp(loc[..., 2]) * pw
    h = np.exp(loc[..., 3]) * ph
    w2 = w / 2
    h2 = h / 2
    return y - h2, x - w2, y + h2, x + w2


def decode_standard(raw, priors, cls_cnt=2, dtype=np.float32):
    """-> [S, 3*g*g, 5+C] rows in concat_bbox order (prior-major, then row, then col) for this scale."""
    raw = raw.astype(dtype)
    S, lh, lw, _ = raw.shape
    d = split(raw, cls_cnt, False)
    y0, x0, y1, x1 = _boxes(d['loc'], priors, lh, lw, dtype)
    rows = np.concatenate([np.stack([y0, x0, y1, x1], -1), sigmoid(d['obj'])[..., None], softmax(d['cls'])], -1)
    return rows.transpose(0, 3, 1, 2, 4).reshape(S, 3 * lh * lw, -1)


def decode_aleatoric(raw, priors, layer_id, cls_cnt=2, dtype=np.float32):
    """-> [S, 3*g*g, 14+C]: [y0,x0,y1,x1, var x4, prod(var), obj, H(obj), cls xC, H(cls), layer_id, prior_id]."""
    raw = raw.astype(dtype)
    S, lh, lw, _ = raw.shape
    d = split(raw, cls_cnt, True)
    y0, x0, y1, x1 = _boxes(d['loc'], priors, lh, lw, dtype)
    loc_var = np.exp(d['log_loc_var'])
    obj = sigmoid(d['obj'])
    cls = softmax(d['cls'])
    ones = np.ones_like(obj)
    rows = np.concatenate([np.stack([y0, x0, y1, x1], -1), loc_var, np.prod(loc_var, -1, keepdims=True),
                           obj[..., None], logistic_entropy(obj)[..., None], cls, softmax_entropy(cls)[..., None],
                           (dtype(layer_id) * ones)[..., None],
                           (np.arange(3, dtype=dtype) * ones)[..., None]], -1)
    return rows.transpose(0, 3, 1, 2, 4).reshape(S, 3 * lh * lw, -1)


def epistemic_stats(raw, cls_cnt=2, dtype=np.float32):
    """decode_epistemic: raw [T,g,g,3*2*(5+C)] -> per-anchor statistics over axis 0 (layers.py:361-411)."""
    raw = raw.astype(dtype)
    d = split(raw, cls_cnt, True)
    loc = d['loc']
    loc_var = np.exp(d['log_loc_var'])
    obj = sigmoid(d['obj'])
    cls = softmax(d['cls'])
    ev_loc = np.mean(loc, axis=0, dtype=dtype)
    ev_loc_locT = np.mean(loc[..., :, None] * loc[..., None, :], axis=0, dtype=dtype)
    epi_covar = ev_loc_locT - ev_loc[..., :, None] * ev_loc[..., None, :]
    obj_mean = np.mean(obj, axis=0, dtype=dtype)
    obj_ent = logistic_entropy(obj_mean)
    obj_mi = obj_ent - np.mean(logistic_entropy(obj), axis=0, dtype=dtype)
    cls_mean = np.mean(cls, axis=0, dtype=dtype)
    cls_ent = softmax_entropy(cls_mean)
    cls_mi = cls_ent - np.mean(softmax_entropy(cls), axis=0, dtype=dtype)
    return {'ev_loc': ev_loc, 'epi_covar_loc': epi_covar, 'ale_var_loc': np.mean(loc_var, axis=0, dtype=dtype),
            'obj_mean': obj_mean, 'obj_mutual_info': obj_mi, 'obj_entropy': obj_ent,
            'cls_mean': cls_mean, 'cls_mutual_info': cls_mi, 'cls_entropy': cls_ent}


def decode_epistemic(raw, priors, layer_id, cls_cnt=2, dtype=np.float32):
    """-> [3*g*g, 21+C] for ONE image: [y0,x0,y1,x1, diag(covar) x4, ale_var x4, det(covar), sum(ale_var),
    obj_mean, obj_MI, obj_H, cls_mean xC, cls_MI, cls_H, layer_id, prior_id]   (layers.py:488-499)."""
    T, lh, lw, _ = raw.shape
    s = epistemic_stats(raw, cls_cnt, dtype)
    y0, x0, y1, x1 = _boxes(s['ev_loc'], priors, lh, lw, dtype)
    cov = s['epi_covar_loc']
    ones = np.ones_like(s['obj_mean'])
    rows = np.concatenate([np.stack([y0, x0, y1, x1], -1),
                           np.diagonal(cov, axis1=-2, axis2=-1), s['ale_var_loc'],
                           np.linalg.det(cov).astype(dtype)[..., None],          # Appendix B-7 (LU)
                           np.sum(s['ale_var_loc'], -1, keepdims=True),
                           s['obj_mean'][..., None], s['obj_mutual_info'][..., None], s['obj_entropy'][..., None],
                           s['cls_mean'], s['cls_mutual_info'][..., None], s['cls_entropy'][..., None],
                           (dtype(layer_id) * ones)[..., None],
                           (np.arange(3, dtype=dtype) * ones)[..., None]], -1)
    return rows.transpose(2, 0, 1, 3).reshape(3 * lh * lw, -1)


def rows_from_raw(variant, raws, priors_by_scale, cls_cnt=2, dtype=np.float32):
    """concat_bbox over the three scales (stride 32, 16, 8).
    standard/aleatoric: raws[j] is [B,g,g,ch] -> [B,N,D];  epistemic: raws[j] is [T,g,g,ch] -> [N,D]."""
    parts = []
    for j, raw in enumerate(raws):
        if variant == 'standard':
            parts.append(decode_standard(raw, priors_by_scale[j], cls_cnt, dtype))
        elif variant == 'aleatoric':
            parts.append(decode_aleatoric(raw, priors_by_scale[j], j, cls_cnt, dtype))
        else:
            parts.append(decode_epistemic(raw, priors_by_scale[j], j, cls_cnt, dtype))
    return np.concatenate(parts, axis=-2)
