"""Generates tests/golden/*.npz by executing the REFERENCE'S OWN Python from /root/reference
(lib_yolo/{layers,model,yolov3}.py graph construction, inference_*.py concat_bbox + nms, and
lib_yolo/utils.py:72-123 numpy decode) on the eager numpy TensorFlow stand-in in oracle/tf_shim.

Run in the build container only (needs /root/reference):   python tests/golden/gen_golden.py
The GPU box never runs this; it only reads the committed .npz files.

Inputs are NOT stored: they are regenerated from seeds by tests/golden_inputs.py (numpy Generator streams are
stable), weights come from byolo.weights.synthetic(variant, seed).  Stored: outputs of the reference code.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, 'oracle', 'tf_shim'))      # fake `tensorflow`, `matplotlib`
sys.path.insert(0, '/root/reference')                            # the reference's lib_yolo / inference_*.py win over the
sys.path.append(os.path.join(ROOT, 'bayesian-yolov3_b200'))      # same-named shims of this repo (only `byolo` is taken from it)
sys.path.insert(1, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))

import tensorflow as tf                                          # noqa: E402  (the shim)
from lib_yolo import yolov3 as ref_yolov3                        # noqa: E402  (reference code)
from lib_yolo import utils as ref_utils                          # noqa: E402
import inference_standard_yolov3 as ref_std                      # noqa: E402
import inference_aleatoric as ref_ale                            # noqa: E402
import inference_epistemic as ref_epi                            # noqa: E402
from byolo import weights as W                                   # noqa: E402
from oracle import philox                                        # noqa: E402
import golden_inputs as GI                                       # noqa: E402

assert tf.__file__.startswith(os.path.join(ROOT, 'oracle', 'tf_shim'))
assert ref_yolov3.__file__.startswith('/root/reference')


class Provider:
    """Serves variables in creation order (conv kernel, then its BN statistics) and Philox dropout masks."""

    def __init__(self, weights, seed, image):
        self.w, self.i, self.d, self.seed, self.image = weights, -1, 0, seed, image

    def next_conv(self, k, cin, cout, use_bias):
        self.i += 1
        w = self.w[self.i]
        assert ('bias' in w) == bool(use_bias)
        return w

    def next_bn(self, c):
        return self.w[self.i]

    def next_dropout_mask(self, shape, rate):
        T, H, Wd, C = shape
        m = np.stack([philox.keep_mask(self.seed, self.d, t, self.image, (H, Wd, C), rate) for t in range(T)])
        self.d += 1
        return m


VARIABLE_NAMES = {}


def config_for(case):
    return {'full_img_size': list(case['img_size']), 'crop': False, 'cls_cnt': case['cls_cnt'],
            'priors': ref_yolov3.ECP_9_PRIORS, 'aleatoric_loss': True, 'inference_mode': True, 'T': case.get('T'),
            'training': False}


def run_case(case, work_dtype=np.float32):
    tf.set_work_dtype(work_dtype)
    variant = case['variant']
    weights = W.synthetic(variant, case['cls_cnt'], case['weight_seed'])
    imgs = GI.images(case).astype(work_dtype)
    cls, mod = {'standard': (ref_yolov3.yolov3, ref_std), 'aleatoric': (ref_yolov3.yolov3_aleatoric, ref_ale),
                'epistemic': (ref_yolov3.bayesian_yolov3_aleatoric, ref_epi)}[variant]
    out = {}
    if variant == 'epistemic':
        rows_all, sel_all, raws_all = [], [], []
        for b in range(imgs.shape[0]):                       # the reference is batch-1 (inference_epistemic.py:193)
            tf.set_provider(Provider(weights, case['dropout_seed'], b))
            tf.reset_names()
            model = cls(config_for(case)).init_model(inputs=tf.Tensor(imgs[b:b + 1]), training=False).get_model()
            rows = mod.concat_bbox([dl.bbox for dl in model.det_layers])
            rows_all.append(rows.a)
            sel_all.append(mod.nms(rows, model).a)
            raws_all.append([dl.raw_output.a for dl in model.det_layers])
            assert (model.obj_idx, model.cls_start_idx) == (14, 17)
        out['rows'] = np.stack(rows_all)
        out['nms_count'] = np.array([len(s) for s in sel_all], np.int32)
        pad = np.zeros((len(sel_all), 1000, rows_all[0].shape[1]), rows_all[0].dtype)
        for b, s in enumerate(sel_all):
            pad[b, :len(s)] = s
        out['nms_rows'] = pad
        for j in range(3):
            out['raw%d' % j] = np.stack([r[j] for r in raws_all])          # [B,T,g,g,42]
    else:
        tf.set_provider(Provider(weights, 0, 0))
        tf.reset_names()
        model = cls(config_for(case)).init_model(inputs=tf.Tensor(imgs), training=False).get_model()
        rows = mod.concat_bbox([dl.bbox for dl in model.det_layers])
        out['rows'] = rows.a
        # the reference's batched nms() tf.concat's per-image results and so only works when every image
        # yields the same count; call it one image at a time and pad to [B,1000,D] + count
        sel = [mod.nms(rows[b:b + 1], model).a[0] for b in range(imgs.shape[0])]
        out['nms_count'] = np.array([len(s) for s in sel], np.int32)
        out['nms_rows'] = np.zeros((len(sel), 1000, rows.a.shape[-1]), rows.a.dtype)
        for b, s_ in enumerate(sel):
            out['nms_rows'][b, :len(s_)] = s_
        for j in range(3):
            out['raw%d' % j] = model.det_layers[j].raw_output.a
    out['dn_out'] = model.dn_out.a
    out['l36'] = model.layers[36].a[:, ::4, ::4, ::8]                      # strided samples keep the file small
    out['l61'] = model.layers[61].a[:, ::2, ::2, ::8]
    VARIABLE_NAMES[variant] = [[n, list(sh)] for n, sh in tf.VARIABLES]      # names the reference's graph code creates
    return out


def main():
    for name, case in GI.CASES.items():
        o32 = run_case(case, np.float32)
        save = {k: v.astype(np.float32) if v.dtype.kind == 'f' else v for k, v in o32.items()}
        if case.get('fp64'):
            o64 = run_case(case, np.float64)
            save['rows64'] = o64['rows']                                   # high-precision rows (float64)
        np.savez_compressed(os.path.join(HERE, name + '.npz'), **save)
        print(name, {k: v.shape for k, v in save.items()}, 'nan:', int(np.isnan(save['rows']).sum()))

    # utils.py:72-123 numpy decode of raw aleatoric head outputs (the only non-TF restatement in the reference)
    rng = np.random.default_rng(7)
    pred = (rng.standard_normal((2, 5, 7, 42)) * 1.5).astype(np.float32)
    pri = ref_yolov3.ECP_9_PRIORS[16]
    got = ref_utils.predictions_to_boxes_numpy_reference_implementation(pred, 2, pri, box_format='yxyx')
    np.savez_compressed(os.path.join(HERE, 'utils_numpy_decode.npz'), pred=pred, boxes=got,
                        priors=np.array([[p.h, p.w] for p in pri]))
    # prior tables, digit for digit
    tabs = {n: np.array([[p.h, p.w] for s in (32, 16, 8) for p in getattr(ref_yolov3, n)[s]])
            for n in ('CITY_PERSONS_9_PRIORS', 'ECP_9_PRIORS', 'ECP_NIGHT_9_PRIORS', 'ECP_DAY_NIGHT_9_PRIORS',
                      'ECP_BIC_9_PRIORS')}
    np.savez_compressed(os.path.join(HERE, 'priors.npz'), **tabs)
    # checkpoint variable names (scope structure from the reference's model.py / yolov3.py, uniquified the TF way)
    import json
    with open(os.path.join(HERE, 'tf_variable_names.json'), 'w') as f:
        json.dump(VARIABLE_NAMES, f, indent=0)
    print('done')


if __name__ == '__main__':
    main()
