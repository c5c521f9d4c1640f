#!/bin/sh
# N=2: gather time with NCCL's default channels per peer and with more of them
for ch in default 8 32; do
  if [ $ch = default ]; then unset NCCL_MIN_P2P_NCHANNELS; else export NCCL_MIN_P2P_NCHANNELS=$ch; fi
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 2 --steps 5 --warmup 3 --no-cpu > gpurun_out/r2q_n2_$ch.json 2> gpurun_out/r2q_n2_$ch.err; echo "ch=$ch rc=$?"
  python - <<PY
import json
d=json.load(open("gpurun_out/r2q_n2_$ch.json"))
print("$ch", "weak", round(d["value"],1), "ms", round(d["ms_per_step"],3), "kernel_ms", round(d["roofline"]["kernel_ms"],3), "strong", round(d["strong"]["value"],1), round(d["strong"]["ms_per_step"],3), d["strong"]["kernel_ms_max_over_ranks"])
PY
done
