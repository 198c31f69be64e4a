#!/bin/bash
# round 2, GPU call V (1 GPU): row variant of the warp-chain kernel (PML-row chunks x ordinary strips) -- parity, then A/B
set -u
out=gpurun_out/r2v; mkdir -p $out
(time timeout 900 python -m pytest tests/test_gpu_fd2d.py -x -q -m gpu -k "warp_chain or bench_launch_plan or deep_passes or baseline_config_5 or smoke" 2>&1 | tail -5) 2>&1 | tail -7
run() { name=$1; shift; env "$@" timeout 600 python bench.py --warmup 5 --no-e2e --no-cpu --no-configs $ARGS > $out/bench_$name.json 2> $out/bench_$name.err
python - <<PY
import json
try:
    d=json.load(open("$out/bench_$name.json"))
    print("%-30s %8.1f Gcell/s  %.4f ms/step  %s" % ("$name", d["value"]/1e3, d["ms_per_step"], d["config"]["pass_depths"][:3]))
except Exception as e:
    print("$name failed", e); print(open("$out/bench_$name.err").read()[-800:])
PY
}
ARGS="--steps 96"; run k96_col_and_row FDTD_COL_FAST=3
ARGS="--steps 96"; run k96_col_only FDTD_COL_FAST=1
ARGS="--steps 20"; run k20_col_and_row FDTD_COL_FAST=3
ARGS="--steps 20"; run k20_col_only FDTD_COL_FAST=1
