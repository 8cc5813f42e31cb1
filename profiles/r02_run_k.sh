#!/bin/bash
# round 2, GPU call K (8 GPUs): the driver's N=8 bench command (config 4 strong scaling + config 2 + config 5 blocks) and the
# end-to-end scaling probe (link ceiling, two transports)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/k_topo.txt 2>&1; nproc > gpurun_out/k_host.txt; free -g >> gpurun_out/k_host.txt; lscpu | grep -E "Model name|Socket|NUMA node\(s\)" >> gpurun_out/k_host.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/k_bench_n8.json 2> gpurun_out/k_bench_n8.err; echo "bench n8 rc=$?"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29522 profiles/e2e_scaling.py > gpurun_out/k_e2e_n8.json 2> gpurun_out/k_e2e_n8.err; echo "e2e n8 rc=$?"
tail -n 3 gpurun_out/k_bench_n8.err gpurun_out/k_e2e_n8.err | cut -c1-300
