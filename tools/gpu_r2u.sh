#!/bin/sh
# tile size: 4 KiB (default, 9 stages at cfg5) vs 2 KiB (17 stages) vs 1 KiB (31 stages); v_old = round-1 ring
WL="cfg5 cfg4 names names-cpw" sh tools/gpu_variants.sh > gpurun_out/r2u_variants.log 2>&1; cat gpurun_out/r2u_variants.log
