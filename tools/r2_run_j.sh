#!/bin/bash
# round 2, GPU call J (1 GPU): the warp-chain interior kernel (fd2d_chain.cu) -- parity, then throughput by shape against
# the shipped deep / register-pipeline passes at 32768^2
set -u
out=gpurun_out/r2j; mkdir -p $out
(time timeout 900 python -m pytest tests/test_gpu_fd2d.py -x -q -m gpu -k "warp_chain" 2>&1 | tail -12) > $out/pytest_chain.txt 2>&1; cat $out/pytest_chain.txt
run() {  # name, env..., args
  name=$1; shift
  env "$@" timeout 600 python bench.py --warmup 5 --no-e2e --no-cpu --no-configs $ARGS > $out/bench_$name.json 2> $out/bench_$name.err
  python - <<PY
import json
try:
    d=json.load(open("$out/bench_$name.json"))
    print("%-28s %8.1f Gcell/s  %.4f ms/step  depths %s" % ("$name", d["value"]/1e3, d["ms_per_step"], d["config"]["pass_depths"][:4]))
except Exception as e:
    print("$name failed", e); print(open("$out/bench_$name.err").read()[-1500:])
PY
}
ARGS="--steps 96 --tblock 8";  run t8_chain_g2k4_w8_ptxsts FDTD_VARIANT=10
ARGS="--steps 96 --tblock 8";  run t8_chain_g2k4_w8_csts FDTD_VARIANT=14
ARGS="--steps 96 --tblock 8";  run t8_chain_g2k4_w12_box2x3_ptxsts FDTD_VARIANT=11
ARGS="--steps 96 --tblock 8";  run t8_chain_g2k4_w12_box2x3_csts FDTD_VARIANT=15
ARGS="--steps 96 --tblock 8";  run t8_chain_g2k4_w12_box3x2 FDTD_VARIANT=12
ARGS="--steps 96 --tblock 12"; run t12_chain_g4k3_w12 FDTD_VARIANT=10
