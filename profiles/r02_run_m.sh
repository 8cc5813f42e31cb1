#!/bin/bash
# round 2, GPU call M: adjacency table without atomics (optimistic placement by local index + count check); tests, bench, A/B
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/m_gpu_tests.log 2>&1; echo "gpu tests rc=$?"
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/m_bench.json 2> gpurun_out/m_bench.err; echo "bench rc=$?"
FEGPU_ADJ_PLACE=0 timeout 600 python bench.py --steps 10 --warmup 3 --no-secondary > gpurun_out/m_bench_atomic.json 2> gpurun_out/m_bench_atomic.err; echo "bench atomic rc=$?"
tail -n 3 gpurun_out/m_gpu_tests.log
