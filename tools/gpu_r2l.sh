#!/bin/sh
# N=8 validation of the bench line (weak, strong, parity, e2e with the merge)
nvidia-smi -L | head -8
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 8 --steps 3 --warmup 3 > gpurun_out/r2l_bench_n8.json 2> gpurun_out/r2l_bench_n8.err; echo "bench n8 rc=$?"; cut -c1-600 gpurun_out/r2l_bench_n8.json; tail -4 gpurun_out/r2l_bench_n8.err
