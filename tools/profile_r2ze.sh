#!/bin/bash
# Round 2, GPU visit ze: the clean build of HEAD — smoke() and the GPU suite, as the driver runs them at round end
set -u
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/gpu_tests_r2ze.log 2>&1; echo "pytest rc=$?" >> gpurun_out/gpu_tests_r2ze.log
tail -4 gpurun_out/gpu_tests_r2ze.log
