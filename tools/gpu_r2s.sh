#!/bin/sh
# N=4: the gather through the IPC window
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 4 --steps 5 --warmup 3 --no-cpu > gpurun_out/r2s_n4.json 2> gpurun_out/r2s_n4.err; echo "rc=$?"; tail -2 gpurun_out/r2s_n4.err | cut -c1-300
python - <<PY
import json
d=json.load(open("gpurun_out/r2s_n4.json"))
print("weak", round(d["value"],1), "ms", round(d["ms_per_step"],3), "kernel_ms", round(d["roofline"]["kernel_ms"],3), "strong", round(d["strong"]["value"],1), round(d["strong"]["ms_per_step"],3), round(d["strong"]["kernel_ms_max_over_ranks"],3), "parity", d["parity"]["ok"], "e2e", round(d["e2e"]["value"],1))
PY
