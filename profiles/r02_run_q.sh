#!/bin/bash
# round 2, GPU call Q: bench.py under the gather pipelining modes (same measurement as every other number of the round)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
for m in 0 3 6 1 4; do
  FEGPU_GATHER_MODE=$m timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/q_bench_mode$m.json 2> gpurun_out/q_bench_mode$m.err; echo "bench mode $m rc=$?"
done
