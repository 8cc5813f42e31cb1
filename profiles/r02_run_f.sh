#!/bin/bash
# round 2, GPU call F: per-rank cost of the 8-way strong split of config 4, each rank played on one GPU (assembly has no collective)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_tile.py -x -q > gpurun_out/f_tile_tests.log 2>&1; echo "tile rc=$?"
REPS=7 timeout 600 python profiles/emulate_rank.py c4s 1 0 > gpurun_out/f_emulate_p1.jsonl 2> gpurun_out/f_emulate_p1.err; echo "p1 rc=$?"
REPS=7 timeout 600 python profiles/emulate_rank.py c4s 8 0 3 7 > gpurun_out/f_emulate_p8.jsonl 2> gpurun_out/f_emulate_p8.err; echo "p8 rc=$?"
FEGPU_TRACE=1 REPS=1 timeout 600 python profiles/emulate_rank.py c4s 8 3 > /dev/null 2> gpurun_out/f_trace_p8.txt
cat gpurun_out/f_emulate_p1.jsonl gpurun_out/f_emulate_p8.jsonl; tail -3 gpurun_out/f_tile_tests.log
