#!/bin/bash
# Round-1 closing evidence run on one B200 (through gpurun): GPU tests, smoke, bench lines, the incumbent CUDA programs,
# the 1D DFT probe.  Everything lands in gpurun_out/final2/.
set -u
out=gpurun_out/final2
mkdir -p $out
timeout 300 python -m pytest tests/test_gpu_fd1d.py -m gpu -q > $out/pytest_gpu_fd1d.txt 2>&1; tail -3 $out/pytest_gpu_fd1d.txt
timeout 900 python -m pytest tests -m gpu -x -q > $out/pytest_gpu.txt 2>&1; tail -3 $out/pytest_gpu.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $out/smoke.txt 2>&1; tail -2 $out/smoke.txt
timeout 400 python bench.py > $out/bench_n1.json 2> $out/bench_n1.err; cut -c1-400 $out/bench_n1.json
timeout 300 python bench.py --impl reference > $out/bench_reference_arm.json 2> $out/bench_reference_arm.err; cut -c1-200 $out/bench_reference_arm.json
timeout 200 bash tools/run_incumbent_cuda.sh > $out/incumbent_cuda.txt 2>&1; cat $out/incumbent_cuda.txt
timeout 200 python tools/probe_1d_dft.py > $out/probe_1d_dft.txt 2>&1; cat $out/probe_1d_dft.txt
timeout 200 python tools/run_reference_benchmarks.py --quick > $out/reference_benchmark_definitions_quick.txt 2>&1; grep -E "Total|config" $out/reference_benchmark_definitions_quick.txt
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $out/nvidia_smi.txt 2>&1
ls -la $out
