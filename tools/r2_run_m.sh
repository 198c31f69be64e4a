#!/bin/bash
# round 2, GPU call M (1 GPU): compute-sanitizer on the new kernels (warp-chain, deep, ring careful) -- memcheck over their
# parity tests, racecheck (shared-memory hazards: cp.async ring, queue slots, mbarrier hand-off) on a few of them
set -u
out=gpurun_out/r2m; mkdir -p $out
export PYTHONDONTWRITEBYTECODE=1
(time timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_gpu_fd2d.py -x -q -m gpu -k "warp_chain_passes or (test_deep_passes and 420)" 2>&1 | tail -15) > $out/memcheck.txt 2>&1; tail -6 $out/memcheck.txt
(time timeout 1500 compute-sanitizer --tool racecheck --error-exitcode 7 python -m pytest tests/test_gpu_fd2d.py -x -q -m gpu -k "(warp_chain_passes and 3_2 and (24-8 or 40-12) and (0- or 11-)) or (test_deep_passes and 420 and 24-8 and 1-3)" 2>&1 | tail -25) > $out/racecheck.txt 2>&1; tail -8 $out/racecheck.txt
