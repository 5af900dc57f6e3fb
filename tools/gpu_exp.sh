#!/bin/bash
# A/B runs of library switches; alternate the variants so that thermal drift does not favour one of them.
mkdir -p gpurun_out
run() { # tag, env...
  local tag=$1; shift
  env "$@" timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-sustained --layers > gpurun_out/exp_$tag.json 2> gpurun_out/exp_$tag.txt
  echo "$tag exit $? $(python -c "import json;d=json.load(open('gpurun_out/exp_$tag.json'));print(d['value'], d['breakdown_ms'])" 2>/dev/null)"
}
run base1 A=1
run cg5_1 BYOLO_CG=5
run base2 A=1
run cg5_2 BYOLO_CG=5
