#!/bin/sh
# ring pipeline with ticket groups taken ahead (batch) and mbarrier-based description waits (desc): each alone, both (default), round-1 ring (v_old)
WL="cfg5 cfg4 names names-cpw" sh tools/gpu_variants.sh > gpurun_out/r2p_variants.log 2>&1; cat gpurun_out/r2p_variants.log
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r2p_tests.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/r2p_tests.log
