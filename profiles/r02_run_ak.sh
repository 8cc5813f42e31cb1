#!/bin/bash
# round 2, GPU call AK: the config-5 block of bench.py with its per-kernel roofline fractions (N = 1)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 600 python bench.py --steps 5 --warmup 3 --config5 --no-secondary > gpurun_out/ak_bench_c5.json 2> gpurun_out/ak_bench_c5.err; echo "bench rc=$?"; tail -n 3 gpurun_out/ak_bench_c5.err
