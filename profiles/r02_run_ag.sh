#!/bin/bash
# round 2, GPU call AG (8 GPUs): column-stencil transport with the column window, A/B at N = 8 on one box
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
for st in 1 0 1; do
  FEGPU_XFER_STENCIL=$st timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 2957$st bench.py --gpus 8 --steps 10 --warmup 3 --no-secondary >> gpurun_out/ag_bench_n8_stencil$st.jsonl 2>> gpurun_out/ag_bench_n8_stencil$st.err; echo "stencil $st rc=$?"
done
