#!/bin/bash
# round 2, GPU call H (8 GPUs): 8- and 3-rank parity of the fused peer-store exchange on hardware (stress cases), then the
# driver's bench launch at N=8 (weak; parity_check + strong split), and the strong-scaling arm at 96 steps
set -u
out=gpurun_out/r2h; mkdir -p $out
nvidia-smi -L > $out/gpus.txt; nvidia-smi topo -m > $out/topo.txt 2>&1
(time FDTD_SLAB_WORLDS=8,3 timeout 1500 python -m pytest tests/test_gpu_slab.py -q -rs -m gpu -k "test_slab_equals_single_device" 2>&1 | tail -80) > $out/pytest_slab_world8_3.txt 2>&1; tail -5 $out/pytest_slab_world8_3.txt
P=29617
(time python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $P bench.py --gpus 8 --steps 20 --warmup 5 > $out/bench_n8_k20.json 2> $out/bench_n8_k20.err) 2>&1 | tail -3
head -c 600 $out/bench_n8_k20.json; echo; tail -3 $out/bench_n8_k20.err
(time python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $((P+1)) bench.py --gpus 8 --steps 96 --warmup 12 --scaling strong --no-e2e --no-configs > $out/bench_n8_k96_strong.json 2> $out/bench_n8_k96_strong.err) 2>&1 | tail -3
head -c 400 $out/bench_n8_k96_strong.json; echo
(time python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port $((P+2)) bench.py --gpus 4 --steps 96 --warmup 12 --scaling strong --no-e2e --no-configs > $out/bench_n4_k96_strong.json 2> $out/bench_n4_k96_strong.err) 2>&1 | tail -3
head -c 400 $out/bench_n4_k96_strong.json; echo
