"""Python face of libbyolo: an `Engine` owns one C handle; torch CUDA tensors are only the buffer type
(data_ptr() in, data_ptr() out) - no torch op runs on the hot path."""
import ctypes as C

import numpy as np
import torch

from . import _lib, priors as _priors, weights as _weights


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else None


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _np_ptr(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


class Engine:
    """variant: 'standard' | 'aleatoric' | 'epistemic'; priors: {32,16,8: [Prior]*3} as in the reference config."""

    def __init__(self, variant, img_hw, cls_cnt=2, priors=None, T=1, max_batch=16, precision='fp16', drop_prob=0.1,
                 standard_test_dropout=False, device=None):
        if not torch.cuda.is_available():
            raise RuntimeError('byolo needs a CUDA device (sm_100a); there is no CPU fallback')
        self.lib = _lib.lib()
        self.device = torch.device('cuda', torch.cuda.current_device() if device is None else device)
        torch.cuda.set_device(self.device)
        torch.cuda.current_stream().synchronize()       # makes sure the primary context exists
        self.variant, self.cls_cnt, self.T = variant, cls_cnt, (T if variant == 'epistemic' else 1)
        self.H, self.W = int(img_hw[0]), int(img_hw[1])
        pri = _priors.as_scale_list(priors if priors is not None else _priors.by_stride('ECP_9_PRIORS'))
        cfg = _lib.Config(variant=_lib.VARIANT_ID[variant], height=self.H, width=self.W, cls_cnt=cls_cnt,
                          max_batch=max_batch, T=self.T, precision=_lib.PRECISION_ID[precision],
                          standard_test_dropout=int(bool(standard_test_dropout)), drop_prob=drop_prob)
        flat = [p for scale in pri for p in scale]
        for i, (h, w) in enumerate(flat):
            cfg.prior_h[i], cfg.prior_w[i] = h, w
        self.h = C.c_void_p()
        _lib.check(self.lib.byolo_create(C.byref(cfg), C.byref(self.h)))
        n, d, o, c = (C.c_int32() for _ in range(4))
        _lib.check(self.lib.byolo_output_shape(self.h, C.byref(n), C.byref(d), C.byref(o), C.byref(c)))
        self.N, self.D, self.obj_idx, self.cls_start_idx = n.value, d.value, o.value, c.value

    def close(self):
        if getattr(self, 'h', None) is not None and self.h.value:
            self.lib.byolo_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------------------------------- weights
    def load_weights(self, weights):
        """weights: list of 75 dicts (byolo.weights) or an already packed BYW1 blob."""
        blob = weights if isinstance(weights, (bytes, bytearray)) else _weights.pack(self.variant, weights, self.cls_cnt)
        buf = (C.c_char * len(blob)).from_buffer_copy(blob)
        _lib.check(self.lib.byolo_load_weights(self.h, C.cast(buf, C.c_void_p), len(blob)))
        return self

    # ------------------------------------------------------------------------------------------- hot path
    def _check_img(self, img):
        assert img.is_cuda and img.dtype == torch.float32 and img.is_contiguous()
        assert tuple(img.shape[1:]) == (self.H, self.W, 3), img.shape
        return img.shape[0]

    def forward(self, img, seed=0, image_index0=0, out=None):
        """img [B,H,W,3] fp32 cuda in [0,1) -> rows [B,N,D] (concat_bbox order)."""
        B = self._check_img(img)
        rows = out if out is not None else torch.empty((B, self.N, self.D), dtype=torch.float32, device=img.device)
        _lib.check(self.lib.byolo_forward(self.h, _ptr(img), B, seed, image_index0, _ptr(rows), _stream()))
        return rows

    def detect(self, img, seed=0, image_index0=0, max_out=1000, iou_thr=0.5, want_rows=False):
        """-> (boxes [B,max_out,D] selection order zero padded, count [B] int32, idx [B,max_out] int32[, rows])."""
        B = self._check_img(img)
        dev = img.device
        boxes = torch.empty((B, max_out, self.D), dtype=torch.float32, device=dev)
        idx = torch.empty((B, max_out), dtype=torch.int32, device=dev)
        cnt = torch.empty((B,), dtype=torch.int32, device=dev)
        rows = torch.empty((B, self.N, self.D), dtype=torch.float32, device=dev) if want_rows else None
        _lib.check(self.lib.byolo_detect(self.h, _ptr(img), B, seed, image_index0, iou_thr, max_out, _ptr(rows),
                                         _ptr(boxes), _ptr(idx), _ptr(cnt), _stream()))
        return (boxes, cnt, idx, rows) if want_rows else (boxes, cnt, idx)

    def detect_packed(self, img, seed=0, image_index0=0, max_out=1000, iou_thr=0.5, out=None):
        """detect() writing ONE fp32 block [B,max_out+1,D]: rows in selection order, zero padded, and in row max_out
        (count, 0, ...) - the message of the multi-GPU all-gather (byolo.dist)."""
        B = self._check_img(img)
        if out is None:
            out = torch.empty((B, max_out + 1, self.D), dtype=torch.float32, device=img.device)
        assert out.is_cuda and out.is_contiguous() and tuple(out.shape) == (B, max_out + 1, self.D)
        _lib.check(self.lib.byolo_detect_packed(self.h, _ptr(img), B, seed, image_index0, iou_thr, max_out, _ptr(out), None,
                                                _stream()))
        return out

    def detect_host(self, img_host, seed=0, image_index0=0, max_out=1000, iou_thr=0.5, out=None):
        """Host numpy/pinned-tensor images in, host numpy results out (copies + sync inside): the sess.run analogue."""
        a = img_host.numpy() if isinstance(img_host, torch.Tensor) else np.ascontiguousarray(img_host, np.float32)
        assert a.dtype == np.float32 and a.shape[1:] == (self.H, self.W, 3)
        B = a.shape[0]
        if out is None:
            out = (np.empty((B, max_out, self.D), np.float32), np.empty((B,), np.int32))
        _lib.check(self.lib.byolo_detect_host(self.h, _np_ptr(a), B, seed, image_index0, iou_thr, max_out,
                                              _np_ptr(out[0]), _np_ptr(out[1]), _stream()))
        return out

    def submit_host(self, img_host, out, slot, seed=0, image_index0=0, max_out=1000, iou_thr=0.5):
        """Pipelined detect_host: enqueue H2D -> detect -> D2H for `slot` (0|1); `out` = (rows [B,max_out,D], count [B])
        host arrays (pinned for real overlap) that wait_host(slot) guarantees to be filled."""
        a = img_host.numpy() if isinstance(img_host, torch.Tensor) else img_host
        assert a.dtype == np.float32 and a.shape[1:] == (self.H, self.W, 3) and a.flags['C_CONTIGUOUS']
        o0 = out[0].numpy() if isinstance(out[0], torch.Tensor) else out[0]
        o1 = out[1].numpy() if isinstance(out[1], torch.Tensor) else out[1]
        assert o0.dtype == np.float32 and o1.dtype == np.int32
        _lib.check(self.lib.byolo_submit_host(self.h, _np_ptr(a), a.shape[0], seed, image_index0, iou_thr, max_out,
                                              _np_ptr(o0), _np_ptr(o1), slot, _stream()))

    def wait_host(self, slot):
        _lib.check(self.lib.byolo_wait_host(self.h, slot))

    def decode(self, raws, B):
        """raws: three dense fp32 cuda tensors [B*T,g,g,ch] -> rows [B,N,D]."""
        rows = torch.empty((B, self.N, self.D), dtype=torch.float32, device=raws[0].device)
        _lib.check(self.lib.byolo_decode(self.h, _ptr(raws[0]), _ptr(raws[1]), _ptr(raws[2]), B, _ptr(rows), _stream()))
        return rows

    def activation(self, conv_index):
        """Dense fp32 [S,H,W,C] copy of the output of conv `conv_index` (0..74) of the last forward."""
        shape = (C.c_int32 * 4)()
        cap = 1 << 28
        # query shape first with a tiny call is not possible without capacity; allocate by known upper bound lazily
        probe = torch.empty((1,), dtype=torch.float32, device=self.device)
        rc = self.lib.byolo_get_activation(self.h, conv_index, _ptr(probe), 0, C.byref(shape), _stream())
        n = int(shape[0]) * int(shape[1]) * int(shape[2]) * int(shape[3])
        assert rc < 0 and n > 0 and n <= cap, (rc, list(shape))
        dst = torch.empty(tuple(int(s) for s in shape), dtype=torch.float32, device=self.device)
        _lib.check(self.lib.byolo_get_activation(self.h, conv_index, _ptr(dst), n, C.byref(shape), _stream()))
        return dst

    def profile(self, enable=True):
        """False/0: off; True/1: one CUDA event before every launch; 2: coarse (stem | conv stack | decode+NMS per call)."""
        _lib.check(self.lib.byolo_profile(self.h, int(enable)))

    def profile_read_coarse(self):
        """Coarse-mode records of the most recent detect() calls (up to 256): array [n, 3] of ms (stem, conv stack, decode+NMS)."""
        cap = 256
        a, b, c = (np.zeros(cap, np.float32) for _ in range(3))
        n = _lib.check(self.lib.byolo_profile_read_coarse(self.h, _np_ptr(a), _np_ptr(b), _np_ptr(c), cap))
        return np.stack([a[:n], b[:n], c[:n]], 1)

    def profile_read(self):
        """Per-launch records of the last profiled detect(): list of dict(ms, kind, layer, flops)."""
        cap = 256
        ms, kind, layer, fl = np.zeros(cap, np.float32), np.zeros(cap, np.int32), np.zeros(cap, np.int32), np.zeros(cap, np.float64)
        mhz = np.zeros(cap, np.float32)
        n = _lib.check(self.lib.byolo_profile_read(self.h, _np_ptr(ms), _np_ptr(kind), _np_ptr(layer), _np_ptr(fl), _np_ptr(mhz), cap))
        names = ('stem', 'conv', 'stack', 'decode', 'nms')
        return [dict(ms=float(ms[i]), kind=names[kind[i]], layer=int(layer[i]), flops=float(fl[i]), sm_mhz=float(mhz[i]))
                for i in range(n)]

    def launch_count(self, B):
        return _lib.check(self.lib.byolo_launch_count(self.h, B))

    def flops_per_image(self, executed=False):
        """Algorithmic FLOPs of the reference graph per image, or (executed=True) what the tensor cores really run."""
        return float((self.lib.byolo_flops_per_image_executed if executed else self.lib.byolo_flops_per_image)(self.h))


def nms(rows, obj_idx, max_out=1000, iou_thr=0.5, packed=False, cluster=0, chunked=False):
    """rows [B,N,D] fp32 cuda -> (boxes [B,max_out,D], count [B], idx [B,max_out]);
    packed=True -> (packed [B,max_out+1,D] whose last row is (count, 0, ...), idx).  cluster / chunked: test hooks."""
    assert rows.is_cuda and rows.dtype == torch.float32 and rows.is_contiguous() and rows.dim() == 3
    B, N, D = rows.shape
    boxes = torch.empty((B, max_out + int(packed), D), dtype=torch.float32, device=rows.device)
    idx = torch.empty((B, max_out), dtype=torch.int32, device=rows.device)
    cnt = None if packed else torch.empty((B,), dtype=torch.int32, device=rows.device)
    with torch.cuda.device(rows.device):
        _lib.check(_lib.lib().byolo_nms_ex(_ptr(rows), B, N, D, obj_idx, iou_thr, max_out, _ptr(boxes), _ptr(idx), _ptr(cnt),
                                           int(packed), cluster, int(chunked), _stream()))
    return (boxes, idx) if packed else (boxes, cnt, idx)


def nms_per_class(rows, obj_idx, cls_start_idx, cls_cnt, max_out=1000, iou_thr=0.5):
    """The per-class NMS of the reference's commented variant (inference_epistemic.py:104-126): for each class, NMS over the
    rows whose score of that class is strictly the largest, results concatenated class by class.  rows [B,N,D] fp32 cuda
    -> list (per image) of numpy arrays [n_b, D], n_b <= cls_cnt * max_out."""
    assert rows.is_cuda and rows.dtype == torch.float32 and rows.is_contiguous() and rows.dim() == 3
    B, N, D = rows.shape
    tmp = torch.empty_like(rows)
    per_image = [[] for _ in range(B)]
    with torch.cuda.device(rows.device):
        for cls in range(cls_cnt):
            _lib.check(_lib.lib().byolo_class_filter(_ptr(rows), B, N, D, obj_idx, cls_start_idx, cls_cnt, cls, _ptr(tmp), _stream()))
            boxes, cnt, _ = nms(tmp, obj_idx, max_out, iou_thr)
            boxes, cnt = boxes.cpu().numpy(), cnt.cpu().numpy()
            for b in range(B):
                sel = boxes[b, :cnt[b]]
                per_image[b].append(sel[sel[:, obj_idx] > -np.inf])       # neutral filler rows (score -inf) come last
    return [np.concatenate(p) for p in per_image]


def conv_layer(x, kernel, bn=None, bias=None, x2=None, residual=None, stride=1, upsample=False, precision='fp16',
               dropout_layer=-1, T=1, seed=0, image_index0=0, drop_prob=0.1, t1=1, t2=1):
    """Per-layer test hook (byolo_conv_layer): x [S,H,W,C1] (+ x2 [S,H,W,C2]) dense fp32 cuda; kernel HWIO numpy;
    bn = dict(beta,gamma,mean,var) or bias array.  Returns dense fp32 [S,Ho,Wo,cout] (x2 size if upsample).
    t1 / t2 > 1: x (alone) / x2 holds S/t samples that the conv reads as S MC-stacked samples (stack_feature_map)."""
    S, H, W, c1 = x.shape
    S *= t1
    c2 = x2.shape[3] if x2 is not None else 0
    k = kernel.shape[0]
    cout = kernel.shape[3]
    kern = np.ascontiguousarray(kernel, np.float32)
    bn_a = np.ascontiguousarray(np.concatenate([bn[n] for n in ('beta', 'gamma', 'mean', 'var')]), np.float32) \
        if bn is not None else None
    bias_a = np.ascontiguousarray(bias, np.float32) if bias is not None else None
    Ho, Wo = H // stride, W // stride
    out = torch.empty((S, 2 * Ho if upsample else Ho, 2 * Wo if upsample else Wo, cout), dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        _lib.check(_lib.lib().byolo_conv_layer(
            _lib.PRECISION_ID[precision], _ptr(x.contiguous()), _ptr(x2.contiguous() if x2 is not None else None), S, H, W,
            c1, c2, k, stride, cout, _np_ptr(kern), _np_ptr(bn_a), _np_ptr(bias_a),
            _ptr(residual.contiguous() if residual is not None else None), int(upsample), dropout_layer, T, seed,
            image_index0, drop_prob, t1, t2, _ptr(out), _stream()))
    return out
