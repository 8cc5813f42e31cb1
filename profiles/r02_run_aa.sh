#!/bin/bash
# round 2, GPU call AA: k_sym_tile with ONE scan of the sorted keys (private rows for the unique lists, per-node copy loops): tests + bench
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tile.py tests/test_gpu_parity.py tests/test_gpu_at_size.py -m gpu -q -x > gpurun_out/aa_tests.log 2>&1; echo "tests rc=$?"; tail -n 2 gpurun_out/aa_tests.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/aa_bench.json 2> gpurun_out/aa_bench.err; echo "bench rc=$?"
