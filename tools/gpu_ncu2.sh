#!/bin/bash
# Single-launch ncu captures (--set full, with source) of selected conv launches (index = layer - 1).
mkdir -p gpurun_out
cap() { # name, kernel regex, skip
  timeout 300 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:$2 -s $3 -c 1 -f -o gpurun_out/prof_$1 python tools/profile_step.py > /dev/null 2>&1
  echo "$1 exit $?"
}
cap l61 conv_umma 60
cap l53 conv_umma 52
