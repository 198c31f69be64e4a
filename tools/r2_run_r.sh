#!/bin/bash
# round 2, GPU call R (1 GPU): the careful kernel as a backfill (interior first, one careful warp per CTA on a default-
# priority stream; FDTD_RING=3) against the shipped order (careful first on a high-priority side stream)
set -u
out=gpurun_out/r2r; mkdir -p $out
(time timeout 600 python -m pytest tests/test_gpu_fd2d.py -x -q -m gpu -k "warp_chain_bench_plan or bench_launch_plan" 2>&1 | tail -3) 2>&1 | tail -4
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
ARGS="--steps 96"; run k96_shipped FDTD_RING=0
ARGS="--steps 96"; run k96_backfill FDTD_RING=3
ARGS="--steps 20"; run k20_shipped FDTD_RING=0
ARGS="--steps 20"; run k20_backfill FDTD_RING=3
ARGS="--steps 96 --tblock 12"; run k96_t12_backfill FDTD_RING=3
ARGS="--steps 96 --size 16384"; run n16384_shipped FDTD_RING=0
ARGS="--steps 96 --size 16384"; run n16384_backfill FDTD_RING=3
