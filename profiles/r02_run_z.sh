#!/bin/bash
# round 2, GPU call Z: two lanes per node in the symbolic kernel (k_sym_pair, FEGPU_SYM_PAIR=1): parity tests, then bench A/B
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
FEGPU_SYM_PAIR=1 timeout 900 python -m pytest tests/test_gpu_tile.py tests/test_gpu_parity.py -m gpu -q -x > gpurun_out/z_tests_pair.log 2>&1; echo "tests pair rc=$?"; tail -n 3 gpurun_out/z_tests_pair.log
for pr in 1 0; do
  FEGPU_SYM_PAIR=$pr timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/z_bench_pair$pr.json 2> gpurun_out/z_bench_pair$pr.err; echo "pair $pr rc=$?"
done
