#!/bin/bash
# A/B runs of library switches (BYOLO_CG, BYOLO_KBS, BYOLO_BN, BYOLO_BSPLIT, BYOLO_NMS_CS ...); alternate the variants so
# that thermal drift does not favour one of them, and compare per-layer cycles (ms x MHz) rather than ms.
mkdir -p gpurun_out
run() { # tag, env...
  local tag=$1; shift
  env "$@" timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-sustained --layers > gpurun_out/exp_$tag.json 2> gpurun_out/exp_$tag.txt
  echo "$tag exit $? $(python -c "import json;d=json.load(open('gpurun_out/exp_$tag.json'));print(d['value'], d['breakdown_ms'])" 2>/dev/null)"
}
run base1 A=1
run var1 BYOLO_CG=1
run base2 A=1
run var2 BYOLO_CG=1
