#!/bin/bash
# round 2, GPU call E (1 GPU): full GPU suite, launch list of the driver's bench command, ncu --set full of the depth-8
# and depth-6 interior kernels and of the ring careful kernel at 32768^2, host vs device pmlparam
set -u
out=gpurun_out/r2e; mkdir -p $out
(time timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -6) > $out/pytest_gpu.txt 2>&1; cat $out/pytest_gpu.txt
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $out/launches_bench_steps20.csv python bench.py --steps 20 --warmup 5 --no-cpu --no-e2e > /dev/null 2>&1; grep -c "k_march\|k_careful" $out/launches_bench_steps20.csv
timeout 600 ncu --set full --import-source on --clock-control none -k regex:"k_march|k_careful2" --launch-skip 6 -c 6 -o $out/prof_bench20 -f python bench.py --steps 20 --warmup 5 --no-cpu --no-e2e --no-configs > $out/ncu_bench20.log 2>&1; tail -2 $out/ncu_bench20.log
python - <<'PY' > $out/pmlparam_host_vs_device.txt 2>&1
import time, numpy as np, torch
from simulation_b200 import fd2d, surface
for nx, ny, npml in ((32768, 32768, 80), (262144, 32768, 80), (262144, 32768, 1000)):
    for where in ("host", "device"):
        fd2d.pmlparam(nx, ny, npml, np.float32, where=where); torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(5):
            fd2d.pmlparam(nx, ny, npml, np.float32, where=where)
        torch.cuda.synchronize()
        print(f"pmlparam {nx}x{ny} npml={npml} where={where}: {(time.perf_counter() - t0) / 5 * 1e3:.3f} ms")
PY
cat $out/pmlparam_host_vs_device.txt
ls -la $out
