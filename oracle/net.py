"""ORACLE (test infrastructure, never shipped): CPU restatement of the reference forward pass.

Restates, op for op, what one `sess.run` of the reference graph computes for the three model classes
(/root/reference/lib_yolo/yolov3.py:232-310 `yolov3`, :370-451 `yolov3_aleatoric`, :518-628
`bayesian_yolov3_aleatoric`) on top of Darknet-53 (/root/reference/lib_yolo/darknet.py:7-39), using torch CPU
ops (fp32 or fp64) for the convolutions.  TensorFlow itself is not installable here; the
primitive-op semantics are the [TF] items of SURVEY.md Appendix B.  The WIRING of this file is pinned against
the reference's own Python (layers.py / model.py / yolov3.py / inference_*.py executed on a numpy TF stand-in,
see oracle/tf_shim and tests/golden/gen_golden.py); the primitive ops are pinned against that stand-in's
independent numpy implementations.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this.
"""
import numpy as np
import torch
import torch.nn.functional as F

from . import philox

BN_EPS = 1e-5          # layers.py:511,516
LEAKY = 0.1            # layers.py:574
DROP_PROB = 0.1        # yolov3.py:462


def conv_specs(variant, cls_cnt=2):
    """The 75 convolutions in creation (= execution = weight-file) order.
    Each: dict(name, k, s, cin, cout, bn, dropout).  `dropout` marks the 15 head convs that have dropout
    before BN in the Bayesian model (yolov3.py:544-550, 575-581, 606-612)."""
    assert variant in ('standard', 'aleatoric', 'epistemic')
    specs = []

    def add(k, s, cin, cout, bn=True, dropout=False, name=''):
        specs.append(dict(name=name, k=k, s=s, cin=cin, cout=cout, bn=bn, dropout=dropout))

    # ---- darknet53 (darknet.py:7-39) ----
    add(3, 1, 3, 32, name='dn0')
    c = 32
    for filters, blocks in ((32, 1), (64, 2), (128, 8), (256, 8), (512, 4)):
        add(3, 2, c, 2 * filters, name='down%d' % (2 * filters))          # make_darknet_downsample_layer
        c = 2 * filters
        for b in range(blocks):                                            # make_darknet_residual_block
            add(1, 1, c, filters, name='res%d_%d_a' % (c, b))
            add(3, 1, filters, c, name='res%d_%d_b' % (c, b))
    det_ch = 3 * (5 + cls_cnt) if variant == 'standard' else 3 * 2 * (5 + cls_cnt)  # layers.py:601,609
    bayes = variant == 'epistemic'
    # ---- det_net_1 .. 3 (yolov3.py:543-622) ----
    for j, (f, cin0) in enumerate(((512, 1024), (256, 768), (128, 384))):
        if j > 0:
            add(1, 1, 2 * f, f, name='det%d_pre' % (j + 1))               # conv 84 / 96: cin = 2f (=512 / 256)
        cin = cin0
        for i in range(3):
            add(1, 1, cin, f, dropout=bayes, name='det%d_%da' % (j + 1, i))
            add(3, 1, f, 2 * f, dropout=bayes and i < 2, name='det%d_%db' % (j + 1, i))
            cin = 2 * f
        add(1, 1, 2 * f, det_ch, bn=False, name='det%d_out' % (j + 1))
    assert len(specs) == 75
    return specs


def _bn_leaky(x, w, dtype):
    gamma, beta, mean, var = (torch.as_tensor(np.asarray(w[k]), dtype=dtype).view(1, -1, 1, 1)
                              for k in ('gamma', 'beta', 'mean', 'var'))
    x = (x - mean) * torch.rsqrt(var + BN_EPS) * gamma + beta          # Appendix B-2
    return torch.maximum(x * LEAKY, x)                                 # Appendix B-3


class Forward:
    """One forward pass; mirrors ModelBuilder's layer list so that route indices can be written exactly as
    in the reference (model.py:76-99)."""

    def __init__(self, variant, weights, cls_cnt=2, dtype=torch.float32, emulate=None, keep_layers=False):
        self.variant = variant
        self.specs = conv_specs(variant, cls_cnt)
        assert len(weights) == len(self.specs)
        self.weights = weights
        self.dtype = dtype
        self.emulate = emulate          # None | 'half' | 'bf16' | 'split': round conv operands like the GPU kernel does
        #                                 ('split' = the fp16x3 mode: every operand a hi + lo fp16 pair, 22 significant bits)
        self.keep_layers = keep_layers
        self._folded = {}

    # --- primitive ops -------------------------------------------------------------------------------
    def _round(self, x):
        if self.emulate == 'half':
            return x.to(torch.float16).to(self.dtype)
        if self.emulate == 'bf16':
            return x.to(torch.bfloat16).to(self.dtype)
        if self.emulate == 'split':
            hi = x.to(torch.float16).to(self.dtype)
            return hi + (x - hi).to(torch.float16).to(self.dtype)
        return x

    def _kernel(self, idx):
        """OIHW torch kernel; in emulate mode BN scale is folded in and the product rounded (as the engine does)."""
        if idx in self._folded:
            return self._folded[idx]
        w = self.weights[idx]
        k = torch.as_tensor(np.asarray(w['kernel']), dtype=self.dtype).permute(3, 2, 0, 1).contiguous()  # HWIO->OIHW
        if self.emulate and self.specs[idx]['bn']:
            scale = torch.as_tensor(np.asarray(w['gamma']), dtype=self.dtype) * torch.rsqrt(
                torch.as_tensor(np.asarray(w['var']), dtype=self.dtype) + BN_EPS)
            k = k * scale.view(-1, 1, 1, 1)
            k = self._round(k)
        elif self.emulate:
            k = self._round(k)
        self._folded[idx] = k
        return k

    def conv(self, x, idx, drop=None):
        """layers.conv (layers.py:545-575): conv2d(no bias) -> [dropout] -> BN -> leaky.
        `drop` = None or (seed, dropout_layer_id, image_index): Philox keep mask per sample t = batch index."""
        sp, w = self.specs[idx], self.weights[idx]
        k = self._kernel(idx)
        if sp['s'] == 2:                                    # darknet_downsample_padding, layers.py:616-635
            x = F.pad(x, (1, 1, 1, 1))
            y = F.conv2d(x, k, stride=2)
        else:                                               # 'SAME', stride 1 (Appendix B-1)
            y = F.conv2d(x, k, padding=sp['k'] // 2)
        if not sp['bn']:                                    # detection conv: bias, linear (layers.py:600-613)
            return y + torch.as_tensor(np.asarray(w['bias']), dtype=self.dtype).view(1, -1, 1, 1)
        if drop is not None and sp['dropout']:
            seed, lid, image = drop
            S, C, H, W = y.shape
            m = np.stack([philox.keep_mask(seed, lid, t, image, (H, W, C), DROP_PROB) for t in range(S)])
            m = torch.from_numpy(m).permute(0, 3, 1, 2).to(self.dtype)
            y = y * (1.0 / (1.0 - DROP_PROB)) * m           # Appendix B-4
        if self.emulate:
            beta = torch.as_tensor(np.asarray(w['beta']), dtype=self.dtype)
            scale = torch.as_tensor(np.asarray(w['gamma']), dtype=self.dtype) * torch.rsqrt(
                torch.as_tensor(np.asarray(w['var']), dtype=self.dtype) + BN_EPS)
            shift = beta - torch.as_tensor(np.asarray(w['mean']), dtype=self.dtype) * scale
            y = y + shift.view(1, -1, 1, 1)
            return torch.maximum(y * LEAKY, y)
        return _bn_leaky(y, w, self.dtype)

    # --- graph -------------------------------------------------------------------------------------
    def run(self, img_nhwc, T=None, seed=0, image_index0=0):
        """img_nhwc: [B,H,W,3] in [0,1).  Returns list (per image for epistemic, else one entry for the batch)
        of dicts {'raw': [raw32, raw16, raw8] as NHWC numpy [S,g,g,det_ch]}, S = T (epistemic) or B."""
        x = torch.as_tensor(np.asarray(img_nhwc), dtype=self.dtype).permute(0, 3, 1, 2)
        x = x.contiguous(memory_format=torch.channels_last)
        if self.variant == 'epistemic':
            assert T is not None
            return [self._run_one(x[b:b + 1], T, seed, image_index0 + b) for b in range(x.shape[0])]
        return [self._run_one(x, None, seed, image_index0)]

    def _run_one(self, x, T, seed, image):
        L = []                       # ModelBuilder.__layers
        fused = {}                   # conv index -> its output as the engine stores it (incl. the fused residual add)
        ci = [0]                     # next conv weight index
        di = [0]                     # next dropout layer id
        bayes = T is not None

        def conv(inp):
            idx = ci[0]
            ci[0] += 1
            drop = None
            if bayes and self.specs[idx]['dropout']:
                drop = (seed, di[0], image)
                di[0] += 1
            out = self.conv(self._round(inp), idx, drop)   # emulate: operands are stored rounded (the image too: stem.cu)
            L.append(out)
            fused[idx] = out
            return out

        # darknet53: conv, then 5 x (downsample, n x residual block)      darknet.py:7-39
        conv(x)
        for blocks in (1, 2, 8, 8, 4):
            conv(L[-1])
            for _ in range(blocks):
                conv(L[-1])
                conv(L[-1])
                L.append(self._round(L[-1] + self._round(L[-3])))   # residual: after the activation, model.py:96-99
                fused[ci[0] - 1] = L[-1]
        assert len(L) == 75
        l36, l61, l74 = L[36], L[61], L[74]
        if bayes:                                             # stack_feature_map, layers.py:595-597
            L.append(l74.expand(T, -1, -1, -1))
        raws = []
        for j, route_src in enumerate((None, l61, l36)):
            if j > 0:
                L.append(L[-3])                               # route([-3]) -> conv "79"/"91" (yolov3.py:263-264)
                conv(L[-1])                                   # 84 / 96
                L.append(F.interpolate(L[-1], scale_factor=2, mode='nearest'))   # upsample, Appendix B-5
                src = route_src.expand(T, -1, -1, -1) if bayes else route_src
                if bayes:
                    L.append(src)                             # stack_feature_map(61|36, T)
                    L.append(torch.cat([L[-2], L[-1]], dim=1))   # route([-2,-1]) = [upsampled, stacked]
                else:
                    L.append(torch.cat([L[-1], src], dim=1))     # route([-1, 61|36])
            for _ in range(6):
                conv(L[-1])
            raw = conv(L[-1])                                 # detection conv (not in darknet numbering as YOLO)
            raws.append(raw.permute(0, 2, 3, 1).contiguous().numpy())
        assert ci[0] == 75
        out = {'raw': raws}
        if self.keep_layers:
            out['layers'] = [t.permute(0, 2, 3, 1).contiguous().numpy() for t in L]
            out['conv_out'] = {i: self._round(t).permute(0, 2, 3, 1).contiguous().numpy() if self.specs[i]['bn']
                               else t.permute(0, 2, 3, 1).contiguous().numpy() for i, t in fused.items()}
        return out
