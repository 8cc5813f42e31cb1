#!/bin/bash
# round 2, GPU call AH: e2e A/B of the column-stencil transport at N = 1 on one box (final build)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
for st in 1 0; do
  FEGPU_XFER_STENCIL=$st timeout 600 python bench.py --steps 10 --warmup 3 --no-secondary > gpurun_out/ah_bench_stencil$st.json 2> gpurun_out/ah_bench_stencil$st.err; echo "stencil $st rc=$?"
done
