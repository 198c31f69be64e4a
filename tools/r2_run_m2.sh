#!/bin/bash
# racecheck of the warp-chain kernel after ordering the consumer's reads before its release (warp barrier)
set -u
out=gpurun_out/r2m; mkdir -p $out
export PYTHONDONTWRITEBYTECODE=1
(time timeout 900 compute-sanitizer --tool racecheck --error-exitcode 7 python -m pytest tests/test_gpu_fd2d.py -x -q -m gpu -k "(warp_chain_passes and 3_2 and 24-8 and (0- or 13-)) or (test_deep_passes and 420 and 24-8 and 1-3)" 2>&1 | cut -c1-260 | tail -40) > $out/racecheck2.txt 2>&1; tail -12 $out/racecheck2.txt
python bench.py --no-e2e --no-cpu --no-configs 2>/dev/null | head -c 300; echo
