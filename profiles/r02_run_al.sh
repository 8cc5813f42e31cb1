#!/bin/bash
# round 2, GPU call AL (2 GPUs): the torchrun bench command with the config-5 block switched on (multi-rank path of its new keys)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29581 bench.py --gpus 2 --steps 5 --warmup 3 --config5 > gpurun_out/al_bench_n2_c5.json 2> gpurun_out/al_bench_n2_c5.err; echo "bench rc=$?"; tail -n 2 gpurun_out/al_bench_n2_c5.err | cut -c1-200
