#!/bin/bash
# round 2, GPU call AI: shared-memory carve-out of the GENERAL gather kernel (k_gather of fegpu_pattern.cu: T10, H20), configs 3 and 5
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
for c in -1 85 70; do
  echo "{\"carveout\": $c}" >> gpurun_out/ai_general_gather_carveout.jsonl
  FEGPU_GATHER_GEN_CARVEOUT=$c timeout 600 python profiles/bench_configs.py c3 c5 >> gpurun_out/ai_general_gather_carveout.jsonl 2>> gpurun_out/ai_general_gather_carveout.err; echo "carveout $c rc=$?"
done
