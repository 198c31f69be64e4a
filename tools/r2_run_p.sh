#!/bin/bash
# round 2, GPU call P (1 GPU): full suite + smoke + the driver's bench (both arms) + the 96-step default after the chunk /
# depth-split policy change (384-row chunks for the deep passes at 32768^2, depth 12 in the split: 20 = 12 + 8)
set -u
out=gpurun_out/r2p; mkdir -p $out
(time timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -6) > $out/pytest_gpu.txt 2>&1; cat $out/pytest_gpu.txt
(time python -c "import __graft_entry__ as g; g.smoke()") > $out/smoke.txt 2>&1; tail -3 $out/smoke.txt
(time python bench.py --impl reference --steps 20 --warmup 5 > $out/bench_ref_k20.json 2> $out/bench_ref_k20.err) 2>&1 | tail -3
(time python bench.py --steps 20 --warmup 5 > $out/bench_k20.json 2> $out/bench_k20.err) 2>&1 | tail -3
head -c 900 $out/bench_k20.json; echo; tail -3 $out/bench_k20.err
(time python bench.py --no-e2e --no-cpu --no-configs > $out/bench_k96.json 2> $out/bench_k96.err) 2>&1 | tail -3
head -c 400 $out/bench_k96.json; echo
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file $out/launches_bench_steps20.csv python bench.py --steps 20 --warmup 5 --no-cpu > /dev/null 2>&1; grep -c "k_march\|k_careful" $out/launches_bench_steps20.csv
