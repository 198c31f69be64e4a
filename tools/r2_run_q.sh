#!/bin/bash
# round 2, GPU call Q (1 GPU): ncu --set full of the shipped depth-12 and depth-8 warp-chain passes at 32768^2 (384-row
# chunks) for the roofline traffic figures; then compute-sanitizer (tools/r2_run_m.sh)
set -u
out=gpurun_out/r2q; mkdir -p $out
for TB in 12 8; do
timeout 600 ncu --set full --import-source on --clock-control none -k regex:"k_march_chain|k_careful2" --launch-skip 2 -c 2 -o $out/prof_chain_t$TB -f python bench.py --steps $((2*TB)) --warmup $TB --tblock $TB --no-cpu --no-e2e --no-configs > $out/ncu_chain_t$TB.log 2>&1; tail -1 $out/ncu_chain_t$TB.log
done
bash tools/r2_run_m.sh
