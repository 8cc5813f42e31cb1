#!/bin/bash
# round 2, GPU call L: k_sym_tile with 64-node tiles (8 CTAs per SM) against 128-node tiles (4 CTAs per SM)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_tile.py -q -x > gpurun_out/l_tests.log 2>&1; echo "tests rc=$?"
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/l_bench_t64.json 2> gpurun_out/l_bench_t64.err; echo "bench t64 rc=$?"
FEGPU_TILE_T=128 timeout 600 python bench.py --steps 10 --warmup 3 --no-secondary > gpurun_out/l_bench_t128.json 2> gpurun_out/l_bench_t128.err; echo "bench t128 rc=$?"
tail -n 2 gpurun_out/l_tests.log
