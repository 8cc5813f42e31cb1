#!/bin/bash
# round 2, GPU call O: value planes + component-major gather for H8 elasticity (config 2); dropped-first candidate keys and the
# hoisted dof0 load in k_sym_tile; full GPU test suite; A/B of the vector planes
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/o_gpu_tests.log 2>&1; echo "gpu tests rc=$?"
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/o_bench.json 2> gpurun_out/o_bench.err; echo "bench rc=$?"
FEGPU_VEC_PLANES=0 timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/o_bench_novecplanes.json 2> gpurun_out/o_bench_novecplanes.err; echo "bench novecplanes rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_gather_tile|k_h8_elastic' -s 2 -c 2 -f -o gpurun_out/r02_c2_planes python profiles/prof_elastic.py 128 > gpurun_out/o_ncu_c2.log 2>&1; echo "ncu c2 rc=$?"
python profiles/ncu_summary.py gpurun_out/r02_c2_planes.ncu-rep > gpurun_out/o_ncu_c2_summary.txt 2>&1
python profiles/lineprof.py gpurun_out/r02_c2_planes.ncu-rep k_gather_tile 40 > gpurun_out/o_lines_gather_c2.txt 2>&1
python profiles/lineprof.py gpurun_out/r02_c2_planes.ncu-rep k_h8_elastic 30 > gpurun_out/o_lines_elastic.txt 2>&1
tail -n 3 gpurun_out/o_gpu_tests.log
