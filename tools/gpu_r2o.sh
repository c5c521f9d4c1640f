#!/bin/sh
# throughput of the per-warp-slot pipeline (constant smem layout) next to the round-1 ring (v_old) and the ring with a 256 ns poll
WL="cfg5 cfg4 names names-cpw" sh tools/gpu_variants.sh > gpurun_out/r2o_variants.log 2>&1; cat gpurun_out/r2o_variants.log
