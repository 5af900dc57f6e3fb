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


def stress_rows(B, seed, N=22743, D=23, obj_idx=14, img=608):
    """NMS stress rows in reference order: cell-centred boxes, 5% exact score ties,
    200 planted clusters of 30 heavily overlapping boxes."""
    rng = np.random.default_rng(seed)
    from byolo import priors as P
    pri = np.array([p for s in P.as_scale_list(P.by_stride('ECP_9_PRIORS')) for p in s])
    rows = np.zeros((B, N, D), np.float32)
    off = 0
    for j, stride in enumerate((32, 16, 8)):
        g = img // stride
        for p in range(3):
            n = g * g
            yy, xx = np.meshgrid(np.arange(g), np.arange(g), indexing='ij')
            cy = (yy.reshape(-1) + 0.5 + rng.uniform(-.5, .5, (B, n))) / g
            cx = (xx.reshape(-1) + 0.5 + rng.uniform(-.5, .5, (B, n))) / g
            h = pri[j * 3 + p, 0] * rng.lognormal(0, .5, (B, n))
            w = pri[j * 3 + p, 1] * rng.lognormal(0, .5, (B, n))
            rows[:, off:off + n, 0] = cy - h / 2
            rows[:, off:off + n, 1] = cx - w / 2
            rows[:, off:off + n, 2] = cy + h / 2
            rows[:, off:off + n, 3] = cx + w / 2
            off += n
    rows[:, :, obj_idx] = 1 / (1 + np.exp(-rng.normal(-2, 2, (B, N))))
    rows[:, :, 4:obj_idx] = rng.random((B, N, obj_idx - 4))
    for b in range(B):
        tie = rng.choice(N, N // 20, replace=False)
        rows[b, tie[: len(tie) // 2], obj_idx] = rows[b, tie[len(tie) // 2: 2 * (len(tie) // 2)], obj_idx]
        for c in rng.choice(N, 200, replace=False):
            members = rng.choice(N, 30, replace=False)
            rows[b, members, :4] = rows[b, c, :4] + rng.normal(0, 0.002, (30, 4)).astype(np.float32)
            rows[b, members, obj_idx] = np.clip(rows[b, c, obj_idx] + rng.normal(0, .05, 30), 1e-4, 1 - 1e-4)
    return rows
