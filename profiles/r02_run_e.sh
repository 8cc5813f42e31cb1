#!/bin/bash
# round 2, GPU call E (2 GPUs): NCCL world-2 test of the device-side gather of row blocks, then the driver's N=2 bench command
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/e_topo.txt 2>&1
timeout 600 python -m pytest tests/test_gpu_multirank_nccl.py -x -q > gpurun_out/e_nccl_tests.log 2>&1; echo "nccl tests rc=$?"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/e_bench_n2.json 2> gpurun_out/e_bench_n2.err; echo "bench n2 rc=$?"
tail -n 5 gpurun_out/e_nccl_tests.log; tail -n 5 gpurun_out/e_bench_n2.err
