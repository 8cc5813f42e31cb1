#!/bin/bash
# round 2, GPU call B: internal element order (ascending smallest node id) + thread-per-node kernels; tile tests, the parity suite,
# then the bench line; A/B partners on the same box: FEGPU_ELEM_ORDER=0 (caller's element order) and FEGPU_TILE_GATHER=0.
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_tile.py -x -q --durations=5 > gpurun_out/b_tile_tests.log 2>&1; echo "tile rc=$?"
timeout 1500 python -m pytest tests -m gpu -q --durations=8 > gpurun_out/b_gpu_tests.log 2>&1; echo "gpu tests rc=$?"
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/b_bench.json 2> gpurun_out/b_bench.err; echo "bench rc=$?"
FEGPU_TILE_GATHER=0 timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/b_bench_oldgather.json 2> gpurun_out/b_bench_oldgather.err; echo "bench oldgather rc=$?"
FEGPU_ELEM_ORDER=0 timeout 600 python bench.py --steps 5 --warmup 3 --no-secondary > gpurun_out/b_bench_noorder.json 2> gpurun_out/b_bench_noorder.err; echo "bench noorder rc=$?"
tail -n 3 gpurun_out/b_tile_tests.log gpurun_out/b_gpu_tests.log
