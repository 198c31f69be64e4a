#!/bin/bash
# Round-1 closing run (second part): full GPU suite on the final build, launch list of the bench command, ncu capture
# of the DFT-carrying 1D kernel.  Everything lands in gpurun_out/final4/.
set -u
out=gpurun_out/final4
mkdir -p $out
timeout 900 python -m pytest tests -m gpu -x -q > $out/pytest_gpu.txt 2>&1; tail -3 $out/pytest_gpu.txt
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/launches_bench_steps24.csv python bench.py --steps 24 --warmup 6 --no-cpu --no-e2e > /dev/null 2>&1; grep -c k_march $out/launches_bench_steps24.csv
timeout 300 ncu --set full --clock-control none -k regex:k1_advance_dft --launch-skip 4 -c 1 -o /tmp/k1dft -f python tools/probe_1d_dft.py 1000000 400 > $out/ncu_k1dft.log 2>&1
python tools/ncu_summary.py /tmp/k1dft.ncu-rep > $out/k1_advance_dft_f32_ncu_full_summary.txt 2>&1; head -30 $out/k1_advance_dft_f32_ncu_full_summary.txt
timeout 400 python bench.py > $out/bench_n1.json 2> $out/bench_n1.err; cut -c1-160 $out/bench_n1.json; grep -o '"e2e": {[^}]*}' $out/bench_n1.json | cut -c1-100
ls -la $out
