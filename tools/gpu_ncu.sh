#!/bin/bash
# ncu evidence for one bench step. Big reports stay in /tmp on the box; CSV summaries + a few single-kernel reports come back.
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches.csv python tools/profile_step.py > gpurun_out/ncu_launches.log 2>&1
echo "launch list exit $?"
timeout 1500 ncu --set full --clock-control none --profile-from-start off -f -o /tmp/prof_all python tools/profile_step.py > gpurun_out/ncu_full.log 2>&1
echo "full exit $?"
ncu -i /tmp/prof_all.ncu-rep --page raw --csv > gpurun_out/prof_all_raw.csv 2>/dev/null
ls -la gpurun_out/prof_all_raw.csv
# single launches with source: conv launch indices (0-based among conv_umma launches) = layer - 1
for spec in "2:l03_c3_304" "52:l53_c3_19" "68:l69_c3_76" "69:l70_pw_76_drop"; do
  skip=${spec%%:*}; name=${spec##*:}
  timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:conv_umma -s $skip -c 1 -f -o gpurun_out/prof_$name python tools/profile_step.py > /dev/null 2>&1
  echo "$name exit $?"
done
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:nms_kernel -c 1 -f -o gpurun_out/prof_nms python tools/profile_step.py > /dev/null 2>&1
du -sh gpurun_out
