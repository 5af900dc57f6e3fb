#!/bin/bash
# ncu evidence for one bench step: launch list (gpu__time_duration) + --set full over every launch of the step (raw CSV).
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches.csv python tools/profile_step.py > gpurun_out/ncu_launches.log 2>&1
echo "launch list exit $?"
timeout 1500 ncu --set full --clock-control none --profile-from-start off -f -o /tmp/prof_all python tools/profile_step.py > gpurun_out/ncu_full.log 2>&1
echo "full exit $?"
ncu -i /tmp/prof_all.ncu-rep --page raw --csv > gpurun_out/prof_all_raw.csv 2>/dev/null
ls -la gpurun_out/prof_all_raw.csv
