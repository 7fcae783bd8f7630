#!/bin/bash
# round 2, call 4 (2 GPUs): the multi-GPU tests that a 1-GPU box skips + the bench line at N=2 (parity + also)
out=gpurun_out; mkdir -p $out
nvidia-smi topo -m > $out/r02d_topo_n2.txt 2>&1
timeout -s KILL 1200 python -m pytest tests -m gpu -q > $out/r02d_tests_n2.log 2>&1; echo "gpu tests rc=$?"; tail -8 $out/r02d_tests_n2.log
timeout -s KILL 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29811 bench.py --gpus 2 --steps 5 --warmup 3 > $out/r02d_bench_n2.json 2> $out/r02d_bench_n2.err; echo "bench rc=$?"; tail -c 800 $out/r02d_bench_n2.err; grep '^{' $out/r02d_bench_n2.json | head -c 5000
