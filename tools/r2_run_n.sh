#!/bin/bash
# round 2, GPU call N (1 GPU): short edge chunks (on / off) and the ring careful kernel at 2-wide vectors (config 4)
set -u
out=gpurun_out/r2n; mkdir -p $out
(time timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -6) > $out/pytest_gpu.txt 2>&1; cat $out/pytest_gpu.txt
show() { python - <<PY
import json
try:
    d=json.load(open("$out/bench_$1.json"))
    c=d.get("configs") or {}
    print("%-26s %8.1f Gcell/s  %.4f ms/step  %s  c4 %s" % ("$1", d["value"]/1e3, d["ms_per_step"], d["config"]["pass_depths"][:3], round(c["c4_tfsf_lossy_4096"]["value"]/1e3,1) if "c4_tfsf_lossy_4096" in c else None))
except Exception as e:
    print("$1 failed", e); print(open("$out/bench_$1.err").read()[-1500:])
PY
}
run() { name=$1; shift; env "$@" timeout 600 python bench.py --warmup 5 --no-e2e --no-cpu $ARGS > $out/bench_$name.json 2> $out/bench_$name.err; show $name; }
ARGS="--steps 20";  run k20_edge1 FDTD_EDGE_CHUNKS=1
ARGS="--steps 20";  run k20_edge0 FDTD_EDGE_CHUNKS=0
ARGS="--steps 96 --no-configs";  run k96_edge1 FDTD_EDGE_CHUNKS=1
ARGS="--steps 96 --no-configs";  run k96_edge0 FDTD_EDGE_CHUNKS=0
ARGS="--steps 96 --no-configs";  run k96_edge1_c256 FDTD_EDGE_CHUNKS=1 FDTD_CHUNK_ROWS=256
ARGS="--steps 20";  run k20_edge1_deep2 FDTD_EDGE_CHUNKS=1 FDTD_DEEP=2
ARGS="--steps 96 --tblock 6 --no-configs";  run k96_t6_edge1 FDTD_EDGE_CHUNKS=1
ARGS="--steps 96 --tblock 6 --no-configs";  run k96_t6_edge1_deep2 FDTD_EDGE_CHUNKS=1 FDTD_DEEP=2
