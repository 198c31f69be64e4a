#!/bin/bash
# round 2, GPU call I (1 GPU): full GPU suite + smoke after the fixes; ncu launch list of the driver's bench command;
# ncu --set full of the depth-8 interior kernel and of the ring careful kernel at 32768^2
set -u
out=gpurun_out/r2i; mkdir -p $out
(time timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -8) > $out/pytest_gpu.txt 2>&1; cat $out/pytest_gpu.txt
(time python -c "import __graft_entry__ as g; g.smoke()") > $out/smoke.txt 2>&1; tail -4 $out/smoke.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file $out/launches_bench_steps20.csv python bench.py --steps 20 --warmup 5 --no-cpu > /dev/null 2>&1; grep -c "k_march\|k_careful" $out/launches_bench_steps20.csv
timeout 600 ncu --set full --import-source on --clock-control none -k regex:"k_march_deep|k_careful2" --launch-skip 2 -c 2 -o $out/prof_deep8 -f python bench.py --steps 8 --warmup 8 --tblock 8 --no-cpu --no-e2e --no-configs > $out/ncu_deep8.log 2>&1; tail -2 $out/ncu_deep8.log
ls -la $out
