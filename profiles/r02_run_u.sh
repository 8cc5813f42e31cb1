#!/bin/bash
# round 2, GPU call U: k_sym_tile capped at 2 (pad 60 KB) / 3 (pad 30 KB) resident CTAs so that k_h8_diffusion (other stream) fits beside it
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
for pad in 0 60 30; do
  FEGPU_SYM_PAD_KB=$pad timeout 600 python bench.py --steps 10 --warmup 3 --no-secondary > gpurun_out/u_bench_pad$pad.json 2> gpurun_out/u_bench_pad$pad.err; echo "pad $pad rc=$?"
  python -c "
import json;d=json.loads(open('gpurun_out/u_bench_pad$pad.json').read().strip().splitlines()[-1]);print('pad $pad ms_per_step',d['ms_per_step'],'cached',d['cached']['ms_per_step'])"
done
