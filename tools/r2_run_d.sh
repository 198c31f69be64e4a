#!/bin/bash
# round 2, GPU call D: full GPU test suite, driver-like bench (both arms), chunk sweep of the deep pass
out=gpurun_out/r2d; mkdir -p $out
free -g | head -2 > $out/host.txt; nproc >> $out/host.txt
(time timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -8) > $out/pytest_gpu.txt 2>&1
cat $out/pytest_gpu.txt
(time python bench.py --impl reference --steps 20 --warmup 5 > $out/bench_ref_k20.json 2> $out/bench_ref_k20.err) 2>&1 | tail -3
head -c 1500 $out/bench_ref_k20.json; echo
(time python bench.py --steps 20 --warmup 5 > $out/bench_k20.json 2> $out/bench_k20.err) 2>&1 | tail -3
head -c 3000 $out/bench_k20.json; echo; tail -3 $out/bench_k20.err
for C in 128 512; do
  FDTD_CHUNK_ROWS=$C python bench.py --steps 96 --warmup 5 --tblock 8 --no-e2e --no-cpu --no-configs 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('chunk $C T=8: %.1f Gcell/s' % (d['value']/1e3))"
done
FDTD_VARIANT=2 python bench.py --steps 96 --warmup 5 --tblock 8 --no-e2e --no-cpu --no-configs 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('late fetch T=8: %.1f Gcell/s' % (d['value']/1e3))"
