#!/bin/bash
# round 2, GPU calls X2 / X4: the driver's bench command at N = 2 and N = 4 (final kernels)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
N=${1:-2}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2954$N bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/x_bench_n$N.json 2> gpurun_out/x_bench_n$N.err; echo "bench n$N rc=$?"
