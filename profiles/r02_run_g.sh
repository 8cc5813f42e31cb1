#!/bin/bash
# round 2, GPU call G: block-plane layout + thread per (node, q, p) gather for vector fields; 3-CTA variant of k_h8_elastic; ncu of the kernels
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x --durations=5 > gpurun_out/g_gpu_tests.log 2>&1; echo "gpu tests rc=$?"
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/g_bench.json 2> gpurun_out/g_bench.err; echo "bench rc=$?"
FEGPU_ELASTIC_CTAS=3 timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/g_bench_el3.json 2> gpurun_out/g_bench_el3.err; echo "bench el3 rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_sym_tile|k_gather_tile|k_adj_table|k_h8_diffusion' -s 4 -c 4 -f -o gpurun_out/r02_c4_tile_v2 python profiles/prof_diffusion.py 256 > gpurun_out/g_ncu_c4.log 2>&1; echo "ncu c4 rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_sym_tile|k_gather_tile|k_h8_elastic' -s 3 -c 3 -f -o gpurun_out/r02_c2_tile_v2 python profiles/prof_elastic.py 128 > gpurun_out/g_ncu_c2.log 2>&1; echo "ncu c2 rc=$?"
tail -n 3 gpurun_out/g_gpu_tests.log
