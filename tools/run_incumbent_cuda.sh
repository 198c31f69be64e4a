#!/bin/bash
# The reference's OWN CUDA programs (fd2d/cuda/test_3_2.cu, test_3_3.cu; built unmodified by `make -C oracle` into
# oracle/_ref/) on this box's B200, next to our library on the same benchmark definitions.  A reported comparison only.
# Each program prints "Total compute time on GPU" and ez[2][0:50]; they check no CUDA errors themselves, so the exit
# status and the printed values (all zeros / nan = the kernels faulted) are shown too.
set -u
cd "$(dirname "$0")/.."
for p in ref_cuda_test_3_2 ref_cuda_test_3_2_pmlfix ref_cuda_test_3_3_pmlfix ref_cuda_test_3_2_8192_pmlfix; do
  exe=oracle/_ref/$p
  [ -x $exe ] || { echo "$p: not built"; continue; }
  echo "== $p"
  t0=$(date +%s%N)
  timeout 60 $exe 2>&1 | awk 'NR<=4 || /rror/ {print} NR==5 {print "..."}'
  echo "exit status: ${PIPESTATUS[0]}, wall $(( ($(date +%s%N) - t0) / 1000000 )) ms"
done
