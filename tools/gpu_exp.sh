#!/bin/bash
# Per-layer times under experiment switches (library built with -DBYOLO_DBG_HOOKS).
mkdir -p gpurun_out
run() { # tag, env...
  local tag=$1; shift
  env "$@" timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --layers > gpurun_out/exp_$tag.json 2> gpurun_out/exp_$tag.txt
  echo "$tag exit $? $(python -c "import json;print(json.load(open('gpurun_out/exp_$tag.json'))['value'])" 2>/dev/null)"
}
run base A=1
run ew16 BYOLO_EW=16
run nosplit BYOLO_BSPLIT=0
run kbs1 BYOLO_KBS=1
