#!/bin/bash
# round 2, GPU call B: ncu --set full of the deep interior kernels (depth 12, 8) and the depth-6 kernel, 16384^2
set -x
out=gpurun_out/r2b; mkdir -p $out
for TB in 12 8 6; do
  K=$TB
  pat="regex:k_march"
  timeout 600 ncu --set full --clock-control none --import-source on -k $pat -s 2 -c 2 -o $out/prof_tb$TB -f \
      python bench.py --size 16384 --steps $K --warmup 3 --tblock $TB --no-e2e --no-cpu --no-configs > $out/ncu_tb$TB.log 2>&1
  tail -3 $out/ncu_tb$TB.log
done
ls -la $out
