#!/bin/bash
# round 2, GPU call O (1 GPU): rows per chunk for the deep (warp-chain) and depth-6 passes, by grid size
set -u
out=gpurun_out/r2o; mkdir -p $out
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
for C in 192 256 384; do ARGS="--steps 96 --tblock 8"; run n32768_t8_c$C FDTD_CHUNK_ROWS=$C; done
for C in 128 192 256; do ARGS="--steps 96 --tblock 6"; run n32768_t6_c$C FDTD_CHUNK_ROWS=$C; done
for C in 256 384; do ARGS="--steps 96 --tblock 12"; run n32768_t12_c$C FDTD_CHUNK_ROWS=$C; done
for N in 16384 8192; do
  for C in 64 128 256; do ARGS="--size $N --steps 96 --tblock 8"; run n${N}_t8_c$C FDTD_CHUNK_ROWS=$C; done
  for C in 64 128; do ARGS="--size $N --steps 96 --tblock 6"; run n${N}_t6_c$C FDTD_CHUNK_ROWS=$C; done
done
ARGS="--size 4096 --steps 192 --tblock 8"; run n4096_t8_c64_v4 FDTD_CHUNK_ROWS=64 FDTD_FORCE_V=4
ARGS="--size 4096 --steps 192 --tblock 6"; run n4096_t6_default FDTD_CHUNK_ROWS=0
