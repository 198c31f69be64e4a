#!/bin/bash
# Why does the end-to-end leg vary between boxes?  Host topology, PCIe rates, and the e2e number with / without the
# NVML CPU affinity.
out=gpurun_out/e2evar; mkdir -p $out
{ nvidia-smi topo -m; lscpu | grep -i -E "numa|socket|model name|^CPU\(s\)"; nvidia-smi -q | grep -i -A6 "GPU Link Info" | head -12; free -g | head -2; } > $out/topology.txt 2>&1
cat $out/topology.txt
timeout 120 python tools/probe_pcie.py 2>&1 | tee $out/pcie_default.txt
for mode in affinity noaffinity; do
  if [ $mode = noaffinity ]; then export FDTD_NO_AFFINITY=1; fi
  timeout 200 python bench.py --no-cpu > $out/bench_$mode.json 2>$out/bench_$mode.err
  echo "$mode: $(grep -o '"e2e": {"value": [0-9.]*' $out/bench_$mode.json)  $(grep -o '"value": [0-9.]*' $out/bench_$mode.json | head -1)"
done
