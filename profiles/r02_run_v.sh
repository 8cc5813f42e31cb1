#!/bin/bash
# round 2, GPU call V: thread-per-element k_adj_place; 256-node tiles of k_sym_tile (A/B); tile + parity tests
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tile.py tests/test_gpu_parity.py -m gpu -q -x > gpurun_out/v_tests.log 2>&1; echo "tests rc=$?"; tail -n 2 gpurun_out/v_tests.log
for t in 128 256; do
  FEGPU_TILE_T=$t timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/v_bench_t$t.json 2> gpurun_out/v_bench_t$t.err; echo "tile $t rc=$?"
done
