#!/bin/bash
# round 2, GPU call I: k_h8_elastic single-pass staging (bench) + ncu of the config-2 kernels
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -q -x -k "elastic" > gpurun_out/i_tests.log 2>&1; echo "tests rc=$?"
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/i_bench.json 2> gpurun_out/i_bench.err; echo "bench rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_gather_tile|k_h8_elastic' -s 2 -c 2 -f -o gpurun_out/r02_c2_tile_v3 python profiles/prof_elastic.py 128 > gpurun_out/i_ncu_c2.log 2>&1; echo "ncu c2 rc=$?"
tail -n 3 gpurun_out/i_tests.log
