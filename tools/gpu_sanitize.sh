#!/bin/bash
# compute-sanitizer memcheck over the small-shape GPU tests (per-layer conv cases, decode, NMS edge cases, one forward).
mkdir -p gpurun_out
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 99 --launch-timeout 0 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -p no:cacheprovider \
  -k "conv_fp16 or conv_fp32 or decode or nms_edge or nms_bit_exact or forward_fp16_tensor_core_path" > gpurun_out/sanitize.log 2>&1
echo "memcheck exit $?"
grep -E "ERROR SUMMARY|Invalid|passed|failed|error" gpurun_out/sanitize.log | tail -15
