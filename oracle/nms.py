"""ORACLE (test infrastructure, never shipped): Python front-ends of the NMS restatement.

`nms()` calls the C restatement (oracle/nms_ref.c, built by oracle/Makefile); `nms_numpy()` is an independent
pure-numpy/Python loop of the same published algorithm for small cases, used to cross-check the C build.
Reference call sites: /root/reference/inference_epistemic.py:99-128, inference_aleatoric.py:104-145,
inference_standard_yolov3.py:104-145.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build():
    so = os.path.join(_HERE, 'liboracle_nms.so')
    src = os.path.join(_HERE, 'nms_ref.c')
    if not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(['make', '-s', '-C', _HERE, 'liboracle_nms.so'])
    return so


def _lib():
    global _LIB
    if _LIB is None:
        _LIB = ctypes.CDLL(build())
        _LIB.byolo_oracle_nms.restype = ctypes.c_int
        _LIB.byolo_oracle_nms.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_float,
                                          ctypes.c_int, ctypes.c_void_p]
        _LIB.byolo_oracle_iou.restype = ctypes.c_float
        _LIB.byolo_oracle_iou.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
    return _LIB


def nms(rows, obj_idx, max_out=1000, iou_thr=0.5):
    """rows [N,D] fp32 -> int32 indices in selection order (<= max_out)."""
    rows = np.ascontiguousarray(rows, dtype=np.float32)
    n, d = rows.shape
    out = np.empty(max_out, dtype=np.int32)
    cnt = _lib().byolo_oracle_nms(rows.ctypes.data, n, d, obj_idx, iou_thr, max_out, out.ctypes.data)
    return out[:cnt].copy()


def nms_gather(rows, obj_idx, max_out=1000, iou_thr=0.5):
    """nms + tf.gather: the rows in selection order (what the reference fetches)."""
    idx = nms(rows, obj_idx, max_out, iou_thr)
    return np.asarray(rows, dtype=np.float32)[idx], idx


def iou_numpy(bi, bj):
    f = np.float32
    ymin_i, xmin_i = min(bi[0], bi[2]), min(bi[1], bi[3])
    ymax_i, xmax_i = max(bi[0], bi[2]), max(bi[1], bi[3])
    ymin_j, xmin_j = min(bj[0], bj[2]), min(bj[1], bj[3])
    ymax_j, xmax_j = max(bj[0], bj[2]), max(bj[1], bj[3])
    area_i = f(f(ymax_i - ymin_i) * f(xmax_i - xmin_i))
    area_j = f(f(ymax_j - ymin_j) * f(xmax_j - xmin_j))
    if area_i <= 0 or area_j <= 0:
        return f(0)
    ih = max(f(min(ymax_i, ymax_j) - max(ymin_i, ymin_j)), f(0))
    iw = max(f(min(xmax_i, xmax_j) - max(xmin_i, xmin_j)), f(0))
    inter = f(ih * iw)
    return f(inter / f(f(area_i + area_j) - inter))


def nms_numpy(rows, obj_idx, max_out=1000, iou_thr=0.5):
    rows = np.asarray(rows, dtype=np.float32)
    order = sorted(range(rows.shape[0]), key=lambda i: (-float(rows[i, obj_idx]), i))
    sel = []
    for i in order:
        if len(sel) >= max_out:
            break
        if all(not (iou_numpy(rows[i, :4], rows[j, :4]) > np.float32(iou_thr)) for j in reversed(sel)):
            sel.append(i)
    return np.asarray(sel, dtype=np.int32)
