#!/bin/sh
# full GPU suite, smoke, then the evidence for cfg5 (bench line, launch list, full capture) with the library as committed
timeout 2400 python -m pytest tests -q -m gpu > gpurun_out/r2z_tests.log 2>&1; echo "tests rc=$?"; tail -4 gpurun_out/r2z_tests.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2z_smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/r2z_smoke.log
sh tools/gpu_profile.sh r2z > gpurun_out/r2z_profile.log 2>&1; echo "profile rc=$?"; tail -c 600 gpurun_out/r2z_profile.log
