#!/bin/bash
# round 2, GPU call A: new thread-per-node kernels -- correctness first (sanitizer on small cases, tile tests, whole GPU suite),
# then the bench line with the tile path and, on the same box, with the general path (FEGPU_TILE=0) for the A/B.
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv > gpurun_out/a_gpu.txt
free -g > gpurun_out/a_host.txt; nproc >> gpurun_out/a_host.txt
timeout 400 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_tile.py -x -q -k "test_tile_path_parity and (0 or 2 or 7 or 8) or slab" > gpurun_out/a_sanitizer.log 2>&1; echo "sanitizer rc=$?" >> gpurun_out/a_sanitizer.log
timeout 600 python -m pytest tests/test_gpu_tile.py -x -q --durations=5 > gpurun_out/a_tile_tests.log 2>&1; echo "rc=$?" >> gpurun_out/a_tile_tests.log
timeout 1500 python -m pytest tests -m gpu -q --durations=12 > gpurun_out/a_gpu_tests.log 2>&1; echo "rc=$?" >> gpurun_out/a_gpu_tests.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/a_bench_tile.json 2> gpurun_out/a_bench_tile.err; echo "bench rc=$?"
FEGPU_TILE=0 timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/a_bench_general.json 2> gpurun_out/a_bench_general.err; echo "bench general rc=$?"
tail -3 gpurun_out/a_sanitizer.log gpurun_out/a_tile_tests.log gpurun_out/a_gpu_tests.log
