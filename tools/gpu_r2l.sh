#!/bin/sh
# N=4 validation of the bench line (weak, strong, parity, e2e with the merge)
nvidia-smi -L | head -8
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 4 --steps 3 --warmup 3 > gpurun_out/r2l_bench_n4.json 2> gpurun_out/r2l_bench_n4.err; echo "bench n4 rc=$?"; cut -c1-600 gpurun_out/r2l_bench_n4.json; tail -4 gpurun_out/r2l_bench_n4.err
timeout 600 python -m pytest tests/test_gpu_multi.py -x -q > gpurun_out/r2l_multi.log 2>&1; echo "multi tests rc=$?"; tail -3 gpurun_out/r2l_multi.log
