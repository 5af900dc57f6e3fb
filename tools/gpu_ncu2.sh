#!/bin/bash
# Single-launch ncu captures (--set full, with source) of the kernels that are not at their roofline yet.
mkdir -p gpurun_out
cap() { # name, kernel regex, skip
  timeout 300 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:$2 -s $3 -c 1 -f -o gpurun_out/prof_$1 python tools/profile_step.py > /dev/null 2>&1
  echo "$1 exit $?"
}
cap nms nms_kernel 0
cap decode decode_kernel 0
cap stem stem_mma 0
cap l01 conv_umma 0
cap l03 conv_umma 2
cap l59_det conv_umma 58
cap l67_det conv_umma 66
cap l62_pw_drop conv_umma 61
du -sh gpurun_out
