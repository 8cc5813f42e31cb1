#!/bin/bash
# round 2, GPU call P (second pass: plain loads again, L2 prefetch two elements ahead): pipelining modes of k_gather_tile (scoreboard sharing found in the SASS control codes), config 4 and config 2
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
for m in 0 1 2 3 4 5 6; do
  FEGPU_GATHER_MODE=$m timeout 300 python profiles/prof_gather_modes.py >> gpurun_out/p_gather_modes.jsonl 2>> gpurun_out/p_gather_modes.err; echo "mode $m rc=$?"
done
cat gpurun_out/p_gather_modes.jsonl

