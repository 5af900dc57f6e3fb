#!/bin/bash
# One GPU session: staged so that a hang in one stage cannot eat the whole box time.  Output -> gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,driver_version,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
run() { # name, timeout, args...
  local name=$1 t=$2; shift 2
  echo "=== $name" | tee -a gpurun_out/summary.txt
  timeout $t python -m pytest tests/test_gpu_parity.py -q -m gpu -p no:cacheprovider "$@" > gpurun_out/$name.log 2>&1
  echo "exit $?" | tee -a gpurun_out/summary.txt
  tail -n 3 gpurun_out/$name.log | tee -a gpurun_out/summary.txt
}
: > gpurun_out/summary.txt
run nms 300 -k "nms"
run decode 200 -k "decode"
run conv_fp32 400 -k "conv_fp32"
run conv_fp16 400 -k "conv_fp16"
run fwd_fp32 600 -k "forward_fp32"
run fwd_fp16 600 -k "forward_fp16"
grep -hE "^(FAILED|ERROR)|Error|error" gpurun_out/*.log | head -60 >> gpurun_out/summary.txt
cat gpurun_out/summary.txt
