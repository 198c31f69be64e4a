set -u
out=gpurun_out/final3; mkdir -p $out
timeout 200 bash tools/run_incumbent_cuda.sh > $out/incumbent_cuda.txt 2>&1; cat $out/incumbent_cuda.txt
timeout 400 python bench.py > $out/bench_n1.json 2> $out/bench_n1.err; cut -c1-200 $out/bench_n1.json; grep -o '"e2e": {[^}]*}' $out/bench_n1.json | cut -c1-160; grep -o '"cpu_baseline": {[^}]*}' $out/bench_n1.json | cut -c1-200
timeout 900 python -m pytest tests -m gpu -x -q > $out/pytest_gpu.txt 2>&1; tail -3 $out/pytest_gpu.txt
