"""Probe of the split-fp16 conv numerics on the device: signed bias / spread of byolo_conv_layer(precision='fp16x3')
against an fp64 evaluation of the same folded fp32 operands, for different activation scales (fp16 subnormal lo parts)
and K depths (accumulation length).  Prints one line per case."""
import os
import sys
import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, 'bayesian-yolov3_b200'), os.path.join(ROOT, 'tests')]
import byolo


def run(c, k, cout, scale, S=2, H=16, W=16, prec='fp16x3', wscale=1.0):
    rng = np.random.default_rng(1)
    x = (rng.standard_normal((S, H, W, c)) * scale).astype(np.float32)
    kern = (rng.standard_normal((k, k, c, cout)) * np.sqrt(2.0 / (k * k * c)) * wscale).astype(np.float32)
    bn = dict(beta=np.zeros(cout, np.float32), gamma=np.ones(cout, np.float32), mean=np.zeros(cout, np.float32),
              var=np.ones(cout, np.float32) - np.float32(1e-5))
    got = byolo.conv_layer(torch.from_numpy(x).cuda(), kern, bn=bn, precision=prec).cpu().numpy().astype(np.float64)
    xin = torch.from_numpy(x).double().permute(0, 3, 1, 2)
    kk = torch.from_numpy(kern).double().permute(3, 2, 0, 1)
    y = torch.nn.functional.conv2d(xin, kk, padding=k // 2)
    y = torch.maximum(y, 0.1 * y).permute(0, 2, 3, 1).numpy()
    d = got - y
    rel_scale = np.abs(y).mean()
    bias = (d * np.sign(y)).mean() / rel_scale
    print('%-7s c=%4d k=%d cout=%4d xscale=%-7g wscale=%-6g  signed bias %+.2e  rms err %.2e  max %.2e  (relative to mean|y| = %.3g)' % (
        prec, c, k, cout, scale, wscale, bias, np.sqrt((d ** 2).mean()) / rel_scale, np.abs(d).max() / rel_scale, rel_scale))


for prec in ('fp16x3', 'fp32'):
    for c, k, cout in ((64, 1, 64), (512, 1, 256), (128, 3, 256), (512, 3, 1024)):
        for scale in (1.0, 0.05, 0.003):
            run(c, k, cout, scale, prec=prec)
    run(512, 3, 1024, 1.0, prec=prec, wscale=0.05)
