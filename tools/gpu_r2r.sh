#!/bin/sh
# N=2: the gather through the IPC window (default) and through ncclSend/ncclRecv (OLM_GATHER_WINDOW=0)
for w in 1 0; do
  export OLM_GATHER_WINDOW=$w
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 2 --steps 5 --warmup 3 --no-cpu > gpurun_out/r2r_n2_w$w.json 2> gpurun_out/r2r_n2_w$w.err; echo "window=$w rc=$?"; tail -2 gpurun_out/r2r_n2_w$w.err | cut -c1-300
  python - <<PY
import json
d=json.load(open("gpurun_out/r2r_n2_w$w.json"))
print("window=$w", "weak", round(d["value"],1), "ms", round(d["ms_per_step"],3), "kernel_ms", round(d["roofline"]["kernel_ms"],3), "strong", round(d["strong"]["value"],1), round(d["strong"]["ms_per_step"],3), round(d["strong"]["kernel_ms_max_over_ranks"],3), "parity", d["parity"]["ok"], "e2e", round(d["e2e"]["value"],1))
PY
done
