#!/bin/bash
# round 2, GPU call F (1 GPU): full GPU suite after the pmlparam fix, smoke, driver-like bench (both arms, 20 steps)
set -u
out=gpurun_out/r2f; mkdir -p $out
(time timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -6) > $out/pytest_gpu.txt 2>&1; cat $out/pytest_gpu.txt
(time python -c "import __graft_entry__ as g; g.smoke()") > $out/smoke.txt 2>&1; tail -3 $out/smoke.txt
(time python bench.py --impl reference --steps 20 --warmup 5 > $out/bench_ref_k20.json 2> $out/bench_ref_k20.err) 2>&1 | tail -3
(time python bench.py --steps 20 --warmup 5 > $out/bench_k20.json 2> $out/bench_k20.err) 2>&1 | tail -3
head -c 1200 $out/bench_k20.json; echo; tail -3 $out/bench_k20.err
