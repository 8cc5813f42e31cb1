#!/bin/bash
# round 2, GPU call Z2: occupancy / tile-size variants of k_sym_pair (FEGPU_SYM_PAIR = 2: 3 CTAs at 80 registers, 3: 256-node tiles)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
for pr in 2 3; do
  FEGPU_SYM_PAIR=$pr timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/z_bench_pair$pr.json 2> gpurun_out/z_bench_pair$pr.err; echo "pair $pr rc=$?"
done
