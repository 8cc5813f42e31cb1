#!/bin/bash
# round 2, GPU call AF: column-stencil transport restricted to the non-empty column window: transport tests + at-size tests + bench
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_at_size.py -m gpu -q -x -k "transport or at_size or full_size or skins" > gpurun_out/af_tests.log 2>&1; echo "tests rc=$?"; tail -n 2 gpurun_out/af_tests.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-secondary > gpurun_out/af_bench.json 2> gpurun_out/af_bench.err; echo "bench rc=$?"
