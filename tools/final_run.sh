#!/bin/bash
# Round-end evidence run on one B200 (called through gpurun): tests, bench lines, launch list, ncu summaries.
set -u
out=gpurun_out/final
mkdir -p $out
timeout 900 python -m pytest tests -m gpu -x -q > $out/pytest_gpu.txt 2>&1; tail -2 $out/pytest_gpu.txt
timeout 300 python bench.py > $out/bench_n1.json 2> $out/bench_n1.err; cut -c1-160 $out/bench_n1.json
timeout 300 python bench.py --impl reference > $out/bench_reference_arm.json 2> $out/bench_reference_arm.err; cut -c1-200 $out/bench_reference_arm.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/launches_bench_steps24.csv python bench.py --steps 24 --warmup 6 --no-cpu --no-e2e > /dev/null 2>&1
timeout 300 ncu --set full --import-source on --clock-control none -k regex:k_march --launch-skip 4 -c 2 -o /tmp/march_final -f python bench.py --steps 32 --warmup 8 --no-cpu --no-e2e > $out/ncu_march.log 2>&1
python tools/ncu_summary.py /tmp/march_final.ncu-rep > $out/march_f32_v4_t6_ncu_full_summary.txt 2>&1
timeout 300 ncu --set full --clock-control none -k regex:k1_advance --launch-skip 8 -c 1 -o /tmp/k1_final -f python tools/probe_1d.py 1000000 640 > $out/ncu_k1.log 2>&1
python tools/ncu_summary.py /tmp/k1_final.ncu-rep > $out/k1_advance_f32_ncu_full_summary.txt 2>&1
timeout 600 python tools/run_reference_benchmarks.py > $out/reference_benchmark_definitions.txt 2>&1
timeout 300 python tools/probe_sizes.py > $out/sizes.txt 2>&1
ls -la $out
