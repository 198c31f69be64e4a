#!/bin/bash
# round 2, GPU call A: parity of the new kernels, then deep passes vs depth 6 at the driver's and the default step counts
set -x
out=gpurun_out/r2a; mkdir -p $out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $out/smi.txt
timeout 900 python -m pytest tests/test_gpu_fd2d.py -x -q -k "deep or bench_launch_plan or positional or more_than_three or advance_matches_oracle" 2>&1 | tail -15 > $out/pytest_new.txt
cat $out/pytest_new.txt
for K in 20 96; do
  for TB in 0 6 8; do
    timeout 600 python bench.py --steps $K --warmup 5 --tblock $TB --no-e2e --no-cpu --no-configs > $out/bench_k${K}_tb${TB}.json 2> $out/bench_k${K}_tb${TB}.err
    python - <<PY
import json
try:
    d=json.load(open("$out/bench_k${K}_tb${TB}.json"))
    print("K=$K TB=$TB value %.1f Gcell/s ms/step %.4f depths %s" % (d["value"]/1e3, d["ms_per_step"], d["config"]["pass_depths"]))
except Exception as e:
    print("K=$K TB=$TB failed", e); print(open("$out/bench_k${K}_tb${TB}.err").read()[-2000:])
PY
  done
done
