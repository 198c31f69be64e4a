#!/bin/bash
# round 2, GPU call G (2 GPUs): N-rank == 1-rank parity on hardware (full case matrix at 2 ranks, both halo modes, the
# missing-neighbour case), then the driver's bench launch at N=2 (weak; carries parity_check and the strong split) and the
# strong-scaling arm, reference arm under torchrun
set -u
out=gpurun_out/r2g; mkdir -p $out
nvidia-smi -L > $out/gpus.txt
(time FDTD_SLAB_WORLDS=2 timeout 1500 python -m pytest tests/test_gpu_slab.py -q -rs -m gpu 2>&1 | tail -45) > $out/pytest_slab_world2.txt 2>&1; tail -5 $out/pytest_slab_world2.txt
P=29517
(time python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $P bench.py --gpus 2 --steps 20 --warmup 5 > $out/bench_n2_k20.json 2> $out/bench_n2_k20.err) 2>&1 | tail -3
head -c 600 $out/bench_n2_k20.json; echo; tail -3 $out/bench_n2_k20.err
(time python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $((P+1)) bench.py --gpus 2 --steps 96 --warmup 12 --scaling strong --no-e2e --no-configs > $out/bench_n2_k96_strong.json 2> $out/bench_n2_k96_strong.err) 2>&1 | tail -3
head -c 400 $out/bench_n2_k96_strong.json; echo
(time python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $((P+2)) bench.py --impl reference --gpus 2 --steps 20 --warmup 5 > $out/bench_ref_n2.json 2> $out/bench_ref_n2.err) 2>&1 | tail -3
head -c 300 $out/bench_ref_n2.json; echo
