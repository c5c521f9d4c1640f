#!/bin/sh
# N=2: multi-GPU tests on real hardware, bench at N=2 (NCCL gather in C, strong + parity blocks), bench N=1
nvidia-smi -L
timeout 900 python -m pytest tests/test_gpu_multi.py -x -q > gpurun_out/r2g_tests.log 2>&1; echo "multi tests rc=$?"; tail -4 gpurun_out/r2g_tests.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/r2g_bench_n2.json 2> gpurun_out/r2g_bench_n2.err; echo "bench n2 rc=$?"; cut -c1-1500 gpurun_out/r2g_bench_n2.json; tail -5 gpurun_out/r2g_bench_n2.err
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/r2g_bench_n1.json 2> gpurun_out/r2g_bench_n1.err; echo "bench n1 rc=$?"; cut -c1-3000 gpurun_out/r2g_bench_n1.json; tail -5 gpurun_out/r2g_bench_n1.err
