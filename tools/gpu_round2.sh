#!/bin/bash
mkdir -p gpurun_out; : > gpurun_out/summary.txt
echo "=== tests" | tee -a gpurun_out/summary.txt
timeout 900 python -m pytest tests -x -q -m gpu -p no:cacheprovider > gpurun_out/tests.log 2>&1; echo "exit $?" | tee -a gpurun_out/summary.txt
tail -n 4 gpurun_out/tests.log | tee -a gpurun_out/summary.txt
echo "=== smoke" | tee -a gpurun_out/summary.txt
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -n 3 | tee -a gpurun_out/summary.txt
echo "=== bench" | tee -a gpurun_out/summary.txt
timeout 900 python bench.py --layers > gpurun_out/bench.json 2> gpurun_out/bench_layers.txt; echo "exit $?" | tee -a gpurun_out/summary.txt
cat gpurun_out/bench.json | tee -a gpurun_out/summary.txt
