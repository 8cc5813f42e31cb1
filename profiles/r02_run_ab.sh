#!/bin/bash
# round 2, GPU call AB: column-stencil codec of the result transport (ids per column + dictionary instead of int32 row indices):
# transport tests, the whole GPU suite (the full-size tests fetch through it), bench A/B
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "transport" > gpurun_out/ab_tests_transport.log 2>&1; echo "transport tests rc=$?"; tail -n 3 gpurun_out/ab_tests_transport.log
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/ab_gpu_tests.log 2>&1; echo "gpu tests rc=$?"; tail -n 3 gpurun_out/ab_gpu_tests.log
for st in 1 0; do
  FEGPU_XFER_STENCIL=$st timeout 600 python bench.py --steps 10 --warmup 3 --no-secondary > gpurun_out/ab_bench_stencil$st.json 2> gpurun_out/ab_bench_stencil$st.err; echo "bench stencil $st rc=$?"
done
