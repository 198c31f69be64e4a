#!/bin/bash
# round 2, GPU call L (1 GPU): the warp-chain kernel as the shipped interior kernel of the deep passes -- ncu --set full of
# one depth-8 pass at 32768^2, the full GPU suite + smoke, the driver's bench command (20 steps) and the 96-step default
set -u
out=gpurun_out/r2l; mkdir -p $out
timeout 600 ncu --set full --import-source on --clock-control none -k regex:"k_march_chain" --launch-skip 1 -c 1 -o $out/prof_chain_g2k4 -f python bench.py --steps 8 --warmup 8 --tblock 8 --no-cpu --no-e2e --no-configs > $out/ncu_chain_g2k4.log 2>&1; tail -2 $out/ncu_chain_g2k4.log
(time timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -8) > $out/pytest_gpu.txt 2>&1; cat $out/pytest_gpu.txt
(time python -c "import __graft_entry__ as g; g.smoke()") > $out/smoke.txt 2>&1; tail -4 $out/smoke.txt
(time python bench.py --steps 20 --warmup 5 > $out/bench_k20.json 2> $out/bench_k20.err) 2>&1 | tail -3
head -c 700 $out/bench_k20.json; echo; tail -3 $out/bench_k20.err
(time python bench.py --no-e2e --no-cpu --no-configs > $out/bench_k96.json 2> $out/bench_k96.err) 2>&1 | tail -3
head -c 400 $out/bench_k96.json; echo
