"""One profiled step of the bench workload (for ncu): warm-up outside the profiler range, one byolo_detect inside."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, 'bayesian-yolov3_b200')]
import byolo  # noqa: E402
from byolo import weights as W  # noqa: E402

B = int(os.environ.get('PROF_B', '16'))
eng = byolo.Engine('epistemic', (608, 608), 2, T=10, max_batch=B, precision=os.environ.get('PROF_PRECISION', 'fp16')).load_weights(W.synthetic('epistemic', 2, 0))
img = torch.from_numpy(np.random.default_rng(1).random((B, 608, 608, 3), dtype=np.float32)).cuda()
for _ in range(2):
    eng.detect(img, seed=1003)
torch.cuda.synchronize()
torch.cuda.profiler.start()
eng.detect(img, seed=1003)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print('done')
