#!/bin/bash
# round 2, GPU call T (4 GPUs): strong scaling (one 32768^2 grid over 4 GPUs, 96 steps) and the driver's weak launch (20 steps)
# after the short edge chunks / column variant; every line carries parity_check
set -u
out=gpurun_out/r2t; mkdir -p $out
P=29717
(time python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port $P bench.py --gpus 4 --steps 96 --warmup 12 --scaling strong --no-e2e --no-configs > $out/bench_n4_k96_strong.json 2> $out/bench_n4_k96_strong.err) 2>&1 | tail -3
head -c 330 $out/bench_n4_k96_strong.json; echo
(time python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port $((P+1)) bench.py --gpus 4 --steps 20 --warmup 5 --no-e2e > $out/bench_n4_k20_weak.json 2> $out/bench_n4_k20_weak.err) 2>&1 | tail -3
head -c 330 $out/bench_n4_k20_weak.json; echo; tail -2 $out/bench_n4_k20_weak.err
