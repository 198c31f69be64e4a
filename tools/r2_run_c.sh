#!/bin/bash
# round 2, GPU call C: deep kernel shapes (which arrays stay in registers) at depth 7 / 8, 32768^2
out=gpurun_out/r2c; mkdir -p $out
timeout 300 python -m pytest tests/test_gpu_fd2d.py -x -q -k "deep or bench_launch_plan" 2>&1 | tail -3
for TB in 8 7; do
for V in 0 1 2 3 4; do
  for K in 96; do
    FDTD_VARIANT=$V timeout 600 python bench.py --steps $K --warmup 5 --tblock $TB --no-e2e --no-cpu --no-configs > $out/bench_v${V}_t${TB}_k$K.json 2> $out/bench_v${V}_t${TB}_k$K.err
    python - <<PY
import json
try:
    d=json.load(open("$out/bench_v${V}_t${TB}_k$K.json"))
    print("variant $V T=$TB K=$K value %.1f Gcell/s ms/step %.4f" % (d["value"]/1e3, d["ms_per_step"]))
except Exception as e:
    print("variant $V T=$TB failed", e); print(open("$out/bench_v${V}_t${TB}_k$K.err").read()[-600:])
PY
  done
done
done
