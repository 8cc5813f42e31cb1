#!/bin/bash
# round 2, GPU call S: defaults after the gather analysis (mode 1, carve-out 85 %): full GPU test suite + bench
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/s_gpu_tests.log 2>&1; echo "gpu tests rc=$?"
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/s_bench.json 2> gpurun_out/s_bench.err; echo "bench rc=$?"
tail -n 3 gpurun_out/s_gpu_tests.log
