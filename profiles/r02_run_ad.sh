#!/bin/bash
# round 2, GPU call AD (8 GPUs): the driver's N=8 bench command with the final kernels (config 4 strong scaling, config 2, config 5 with
# the inertial-bisection owner map) + the NCCL world-2 test of the device-side block gather
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/ad_bench_n8.json 2> gpurun_out/ad_bench_n8.err; echo "bench n8 rc=$?"
timeout 600 python -m pytest tests/test_gpu_multirank_nccl.py -m gpu -q > gpurun_out/ad_nccl_tests.log 2>&1; echo "nccl tests rc=$?"; tail -n 2 gpurun_out/ad_nccl_tests.log
tail -n 3 gpurun_out/ad_bench_n8.err | cut -c1-300
