"""A/B experiments: bench.py on the experiment build of the library (make -C bayesian-yolov3_b200/csrc dbg ->
libbyolo_dbg.so, compiled with -DBYOLO_DBG_HOOKS so that the BYOLO_* environment switches are live).
  BYOLO_CHUNK=2 python tools/exp_bench.py --precision fp16x3 --no-cpu-baseline ..."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, 'bayesian-yolov3_b200')]
from byolo import _lib

_lib.LIB_PATH = os.path.join(os.path.dirname(_lib.LIB_PATH), 'libbyolo_dbg.so')
import bench

bench.main()
