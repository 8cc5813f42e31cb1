#!/bin/bash
# round 2, GPU call H: conflict-free accumulator image for vector fields + software-pipelined gather; tests, bench
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x --durations=5 > gpurun_out/h_gpu_tests.log 2>&1; echo "gpu tests rc=$?"
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/h_bench.json 2> gpurun_out/h_bench.err; echo "bench rc=$?"
tail -n 3 gpurun_out/h_gpu_tests.log
