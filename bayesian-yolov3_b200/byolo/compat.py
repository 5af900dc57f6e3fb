"""Minimal stand-ins for the TensorFlow-1.x objects that sit ON the reference's call surface (placeholders, fetch
handles, Session.run, the one-shot dataset iterator), so that detect.py / inference_*.py keep their shape while the
work is done by libbyolo.  Nothing here computes: Session.run resolves a fetch handle to one Engine call.

Reference surface mirrored (paths under /root/reference):
  tf.placeholder / feed_dict / sess.run      detect.py:93,124   inference_epistemic.py:57-76
  tf.errors.OutOfRangeError (end of data)    inference_epistemic.py:69-70
  dataset_utils.TestingDataset(...).iterator.get_next() -> (img, filename)   dataset_utils.py:188-219
"""
import glob
import os

import numpy as np


class OutOfRangeError(Exception):
    """End of the input sequence (tf.errors.OutOfRangeError)."""


class errors:                                             # `tf.errors.OutOfRangeError` spelling
    OutOfRangeError = OutOfRangeError


class _Shape(tuple):
    def as_list(self):
        return list(self)


class Placeholder:
    """tf.placeholder(tf.float32, shape): a named slot that feed_dict fills with a [B,H,W,3] float32 array in [0,1)."""

    def __init__(self, shape, name='img'):
        self.shape, self.name = _Shape(shape), name

    def get_shape(self):
        return self.shape


class IteratorInput:
    """dataset.iterator.get_next()[0]: every Session.run pulls the next batch (images, filenames)."""

    def __init__(self, dataset, shape):
        self.dataset, self.shape = dataset, _Shape(shape)
        self.last_files = None

    def get_shape(self):
        return self.shape

    def next_batch(self):
        imgs, files = self.dataset.next_batch()
        self.last_files = files
        return imgs


class Filenames:
    """dataset.iterator.get_next()[1]: fetch handle for the filenames of the batch the same run consumed."""

    def __init__(self, source):
        self.source = source


class Op:
    """Fetch handle; supports the `op[0, ...]` the reference applies to batched results (detect.py:19,26)."""

    def __getitem__(self, idx):
        return IndexedOp(self, idx)


class IndexedOp(Op):
    def __init__(self, base, idx):
        self.base, self.idx = base, idx


class RowsOp(Op):
    """concat_bbox(...): all candidate rows, [B,N,D] (standard/aleatoric) or [N,D] (epistemic, batch 1)."""

    def __init__(self, model):
        self.model = model


class NmsOp(Op):
    """nms(rows, model): rows kept by class-agnostic NMS(max 1000, IoU 0.5) in selection order.
    epistemic: [n<=1000, D] for the single image.  standard/aleatoric: [B,1000,D]; the reference can only form this
    array when every image keeps the same number of boxes (tf.concat in its while_loop) - here images
    that keep fewer are zero padded and the true counts are on `Session.last_counts`."""

    def __init__(self, rows_op, model, per_class=False):
        self.rows_op, self.model, self.per_class = rows_op, model, per_class


class Session:
    """sess.run(fetches, feed_dict): one run = at most one forward pass per model, shared by all fetches."""

    def __init__(self, config=None, seed=0):
        self.seed = seed
        self.last_counts = None
        self._run_index = 0

    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False

    def run(self, fetches, feed_dict=None):
        single = not isinstance(fetches, (list, tuple))
        fl = [fetches] if single else list(fetches)
        cache = {}
        out = [self._eval(f, feed_dict or {}, cache) for f in fl]
        self._run_index += 1
        return out[0] if single else out

    # --------------------------------------------------------------------------------------------------------
    def _forward(self, model, feed, cache):
        key = id(model)
        if key not in cache:
            src = model.inputs
            if isinstance(src, Placeholder):
                if src not in feed:
                    raise KeyError('placeholder %r was not fed' % src.name)
                img = np.ascontiguousarray(feed[src], np.float32)
            else:
                img = src.next_batch()
            cache[key] = model.execute(img, seed=self.seed + self._run_index)
        return cache[key]

    def _eval(self, f, feed, cache):
        if isinstance(f, IndexedOp):
            return self._eval(f.base, feed, cache)[f.idx]
        if isinstance(f, Filenames):
            if f.source.last_files is None:
                raise RuntimeError('fetch the images of the batch in the same run as the filenames')
            return f.source.last_files
        if isinstance(f, NmsOp):
            res = self._forward(f.model, feed, cache)
            if f.per_class:        # optional variant of inference_epistemic.py:104-126: one NMS per class, concatenated
                import torch
                from . import engine as _engine
                m = f.model
                per = _engine.nms_per_class(torch.from_numpy(res['rows']).cuda(), m.obj_idx, m.cls_start_idx, m.cls_cnt)
                self.last_counts = [len(p) for p in per]
                if m.variant == 'epistemic':
                    return per[0]
                out = np.zeros((len(per), max(self.last_counts + [1]), res['rows'].shape[-1]), np.float32)
                for b, p in enumerate(per):
                    out[b, :len(p)] = p
                return out
            self.last_counts = res['count']
            if f.model.variant == 'epistemic':
                return res['boxes'][0, :res['count'][0]].copy()
            return res['boxes'].copy()
        if isinstance(f, RowsOp):
            res = self._forward(f.model, feed, cache)
            return res['rows'][0].copy() if f.model.variant == 'epistemic' else res['rows'].copy()
        raise TypeError('cannot fetch %r' % (f,))


class ImageDataset:
    """Stand-in for dataset_utils.TestingDataset (TFRecord input is out of scope): iterates over the
    image files matched by config['data']['file_pattern'] (.png/.jpg via cv2 -> RGB float32 [0,1), or .npy arrays
    already in that form), batch_size at a time, once."""

    def __init__(self, config, config_key='data'):
        self.files = sorted(glob.glob(os.path.expandvars(config[config_key]['file_pattern'])))
        self.batch_size = int(config['batch_size'])
        self.size = tuple(config['full_img_size'])
        self.pos = 0
        self.iterator = self

    def get_next(self):
        src = IteratorInput(self, (None,) + self.size)
        return src, Filenames(src)

    def _load(self, path):
        if path.endswith('.npy'):
            img = np.load(path).astype(np.float32)
        else:
            import cv2
            bgr = cv2.imread(path, cv2.IMREAD_COLOR)
            if bgr is None:
                raise IOError('cannot read %s' % path)
            img = bgr[:, :, ::-1].astype(np.float32) / np.float32(255.0)      # convert_image_dtype: uint8 -> [0,1]
        assert img.shape == self.size, '%s has shape %s, config says %s' % (path, img.shape, self.size)
        return img

    def next_batch(self):
        if self.pos >= len(self.files):
            raise OutOfRangeError()
        chunk = self.files[self.pos:self.pos + self.batch_size]
        self.pos += len(chunk)
        imgs = np.stack([self._load(p) for p in chunk])
        names = np.array([[p.encode('utf-8')] for p in chunk], dtype=object)   # files[i][0].decode('utf-8') as in the reference
        return imgs, names
