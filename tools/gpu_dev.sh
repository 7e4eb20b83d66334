#!/bin/bash
# Development GPU session: tests (all failures shown), smoke, short bench, attention micro-benchmarks.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit=$?" | tee gpurun_out/status.txt
tail -n 15 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit=$?" | tee -a gpurun_out/status.txt
timeout 600 python bench.py --steps 10 --warmup 3 --detail gpurun_out/bench_detail.json > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench exit=$?" | tee -a gpurun_out/status.txt
for what in "$@"; do
  timeout 300 python tools/microbench.py $what 20 > gpurun_out/mb_$what.log 2>&1; echo "microbench $what exit=$?" | tee -a gpurun_out/status.txt
  cat gpurun_out/mb_$what.log | tail -n 40
done
cat gpurun_out/status.txt; tail -c 1500 gpurun_out/bench.log; tail -n 5 gpurun_out/bench.err
