"""Seeded inputs shared by tests/golden/gen_golden.py (which stores the reference code's outputs) and the
parity tests (which regenerate the same inputs on any machine)."""
import numpy as np

CASES = {
    # 128x128 -> grids 4/8/16 -> 1008 candidate rows: the 1000-box cap of the reference NMS is hit
    'standard_128': dict(variant='standard', img_size=(128, 128, 3), batch=2, cls_cnt=2, weight_seed=11, img_seed=101),
    'aleatoric_128': dict(variant='aleatoric', img_size=(128, 128, 3), batch=2, cls_cnt=2, weight_seed=12,
                          img_seed=102),
    # non-square 96x160 -> grids 3x5 / 6x10 / 12x20 -> 945 rows; T=6 MC samples (> 4, so the 4x4 covariance has
    # full rank and its determinant column is not pure round-off), 2 images
    'epistemic_96x160': dict(variant='epistemic', img_size=(96, 160, 3), batch=2, cls_cnt=2, weight_seed=13,
                             img_seed=103, T=6, dropout_seed=1003, fp64=True),
}


def images(case):
    rng = np.random.default_rng(case['img_seed'])
    return rng.random((case['batch'],) + tuple(case['img_size']), dtype=np.float32)      # U[0,1) like decode_img
