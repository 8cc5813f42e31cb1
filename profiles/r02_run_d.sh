#!/bin/bash
# round 2, GPU call D: plane (struct-of-arrays) layouts for the thread-per-node kernels; single host round trip per build
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_tile.py -x -q --durations=5 > gpurun_out/d_tile_tests.log 2>&1; echo "tile rc=$?"
timeout 1500 python -m pytest tests -m gpu -q --durations=8 > gpurun_out/d_gpu_tests.log 2>&1; echo "gpu tests rc=$?"
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/d_bench.json 2> gpurun_out/d_bench.err; echo "bench rc=$?"
FEGPU_PLANES=0 timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/d_bench_noplanes.json 2> gpurun_out/d_bench_noplanes.err; echo "bench noplanes rc=$?"
tail -n 3 gpurun_out/d_tile_tests.log gpurun_out/d_gpu_tests.log
