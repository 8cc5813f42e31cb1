#!/bin/bash
# round 2, GPU call C: ncu --set full of the thread-per-node kernels on config 4 (256^3 diffusion) and config 2 (128^3 elasticity)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_sym_tile|k_gather_tile|k_adj_table' -s 3 -c 3 -f -o gpurun_out/r02_c4_tile python profiles/prof_diffusion.py 256 > gpurun_out/c_ncu_c4.log 2>&1; echo "ncu c4 rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_sym_tile|k_gather_tile|k_adj_table' -s 3 -c 3 -f -o gpurun_out/r02_c2_tile python profiles/prof_elastic.py 128 > gpurun_out/c_ncu_c2.log 2>&1; echo "ncu c2 rc=$?"
ls -la gpurun_out/*.ncu-rep
