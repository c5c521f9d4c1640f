#!/bin/sh
# N=8: the bench line (weak, strong, parity, e2e with the merge), records gathered through the IPC window
nvidia-smi -L | wc -l
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 8 --steps 5 --warmup 3 --no-cpu > gpurun_out/r2v_n8.json 2> gpurun_out/r2v_n8.err; echo "rc=$?"; tail -3 gpurun_out/r2v_n8.err | cut -c1-300
python - <<PY
import json
d=json.load(open("gpurun_out/r2v_n8.json"))
print("weak", round(d["value"],1), "ms", round(d["ms_per_step"],3), "kernel_ms", round(d["roofline"]["kernel_ms"],3), "strong", round(d["strong"]["value"],1), round(d["strong"]["ms_per_step"],3), round(d["strong"]["kernel_ms_max_over_ranks"],3), "parity", d["parity"]["ok"], "e2e", round(d["e2e"]["value"],1), d["e2e"].get("h2d_ms_max_over_ranks"))
PY
