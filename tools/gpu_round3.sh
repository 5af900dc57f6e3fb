#!/bin/bash
# Session 2, call 1: confirm the restored tree (tests, smoke, bench), then find what bounds the small backbone layers:
# per-layer times with parts of the conv kernel switched off (BYOLO_DBG) and ncu captures of single launches.
mkdir -p gpurun_out; : > gpurun_out/summary.txt
echo "=== tests" | tee -a gpurun_out/summary.txt
timeout 700 python -m pytest tests -q -m gpu -p no:cacheprovider > gpurun_out/tests.log 2>&1; echo "exit $?" | tee -a gpurun_out/summary.txt
tail -n 5 gpurun_out/tests.log | tee -a gpurun_out/summary.txt
grep -E "^E   .*(Failed|Assertion)" gpurun_out/tests.log | cut -c1-400 | head -10 | tee -a gpurun_out/summary.txt
echo "=== smoke" | tee -a gpurun_out/summary.txt
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -n 3 | tee -a gpurun_out/summary.txt
echo "=== bench" | tee -a gpurun_out/summary.txt
timeout 600 python bench.py --layers > gpurun_out/bench.json 2> gpurun_out/bench_layers.txt; echo "exit $?" | tee -a gpurun_out/summary.txt
cat gpurun_out/bench.json | tee -a gpurun_out/summary.txt
for d in 1 2 4 3; do
  BYOLO_DBG=$d timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --layers > gpurun_out/bench_dbg$d.json 2> gpurun_out/layers_dbg$d.txt
  echo "dbg $d exit $?" | tee -a gpurun_out/summary.txt
done
for spec in "0:l01_s2_304" "1:l02_pw_304" "2:l03_c3_304" "9:l10_pw_76" "69:l70_pw_76_drop"; do
  skip=${spec%%:*}; name=${spec##*:}
  timeout 300 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:conv_umma -s $skip -c 1 -f -o gpurun_out/prof_$name python tools/profile_step.py > /dev/null 2>&1
  echo "$name exit $?" | tee -a gpurun_out/summary.txt
done
du -sh gpurun_out | tee -a gpurun_out/summary.txt
