#!/bin/sh
# per-warp chunk slots: parity suite, then throughput next to the previous library (variants/v_old.so) on the same box
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r2m_tests.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/r2m_tests.log
WL="cfg5 cfg4 names names-cpw" sh tools/gpu_variants.sh > gpurun_out/r2m_variants.log 2>&1; cat gpurun_out/r2m_variants.log
