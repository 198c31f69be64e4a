#!/bin/bash
# round 2, GPU call K (1 GPU): ncu --set full of the warp-chain kernel (G=2, K=4), one depth-8 pass at 32768^2
set -u
out=gpurun_out/r2k; mkdir -p $out
FDTD_VARIANT=10 timeout 600 ncu --set full --import-source on --clock-control none -k regex:"k_march_chain" --launch-skip 1 -c 1 -o $out/prof_chain_g2k4 -f python bench.py --steps 8 --warmup 8 --tblock 8 --no-cpu --no-e2e --no-configs > $out/ncu_chain_g2k4.log 2>&1; tail -2 $out/ncu_chain_g2k4.log
