#!/bin/bash
# Standard GPU session: tests, smoke, bench, then ncu launch list + full captures of the two top kernels.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit=$?" | tee gpurun_out/status.txt
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit=$?" | tee -a gpurun_out/status.txt
timeout 900 python bench.py --steps 10 --warmup 3 --detail gpurun_out/bench_detail.json > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench exit=$?" | tee -a gpurun_out/status.txt
timeout 600 python bench.py --impl reference --steps 4 --warmup 1 > gpurun_out/bench_ref.log 2>&1; echo "bench_ref exit=$?" | tee -a gpurun_out/status.txt
if [ "$1" != "noncu" ]; then
timeout 300 python tools/profile_step.py 32 1 > gpurun_out/profile_plain.log 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python tools/profile_step.py 32 1 > gpurun_out/ncu_launches.log 2>&1
echo "ncu launches exit=$?" | tee -a gpurun_out/status.txt
timeout 300 python tools/profile_step.py 32 1 > gpurun_out/profile_plain2.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:window_attn -s 12 -c 2 -o gpurun_out/prof_attn -f python tools/profile_step.py 32 1 > gpurun_out/ncu_attn.log 2>&1
echo "ncu attn exit=$?" | tee -a gpurun_out/status.txt
timeout 300 python tools/profile_step.py 32 1 > gpurun_out/profile_plain3.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:linear_tc -s 49 -c 4 -o gpurun_out/prof_linear -f python tools/profile_step.py 32 1 > gpurun_out/ncu_linear.log 2>&1
echo "ncu linear exit=$?" | tee -a gpurun_out/status.txt
timeout 300 python tools/profile_step.py 32 1 > gpurun_out/profile_plain4.log 2>&1 &&
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --print-kernel-base demangled -c 600 --csv --log-file gpurun_out/traffic.csv python tools/profile_step.py 32 1 > gpurun_out/ncu_traffic.log 2>&1
echo "ncu traffic exit=$?" | tee -a gpurun_out/status.txt
fi
cat gpurun_out/status.txt; tail -c 3000 gpurun_out/bench.log; tail -n 5 gpurun_out/bench.err; tail -n 3 gpurun_out/pytest_gpu.log; cat gpurun_out/bench_ref.log | tail -n 2
