#!/bin/bash
# round 2, GPU call R: shared-memory carve-out of k_gather_tile (L1 = 256 KB - carve-out holds the lines of the loads in flight) x pipelining mode
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
for c in 40 55 70 85 100; do
  for m in 0 3 6 1 4; do
    echo "{\"carveout\": $c}" >> gpurun_out/r_gather_carveout.jsonl
    FEGPU_GATHER_CARVEOUT=$c FEGPU_GATHER_MODE=$m timeout 300 python profiles/prof_gather_modes.py >> gpurun_out/r_gather_carveout.jsonl 2>> gpurun_out/r_gather_carveout.err
  done
done
cat gpurun_out/r_gather_carveout.jsonl
