#!/bin/bash
# round 2, GPU call K (1 GPU): ncu --set full of the warp-chain kernel shapes (16384^2, one pass each)
set -u
out=gpurun_out/r2k; mkdir -p $out
for TB in 8 12; do
FDTD_VARIANT=10 timeout 600 ncu --set full --import-source on --clock-control none -k regex:"k_march_chain" --launch-skip 1 -c 1 -o $out/prof_chain_t$TB -f python bench.py --size 16384 --steps $TB --warmup $TB --tblock $TB --no-cpu --no-e2e --no-configs > $out/ncu_chain_t$TB.log 2>&1; tail -2 $out/ncu_chain_t$TB.log
done
ls -la $out
