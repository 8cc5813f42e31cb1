#!/bin/bash
# round 2, GPU call T: outer-product formulation of the elasticity kernels for cubic-symmetry (isotropic) D: full GPU suite,
# bench with the config-5 block, A/B with the shortcut off
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/t_gpu_tests.log 2>&1; echo "gpu tests rc=$?"
timeout 900 python bench.py --steps 10 --warmup 3 --config5 > gpurun_out/t_bench.json 2> gpurun_out/t_bench.err; echo "bench rc=$?"
FEGPU_ELASTIC_ISO=0 timeout 900 python bench.py --steps 10 --warmup 3 --config5 > gpurun_out/t_bench_noiso.json 2> gpurun_out/t_bench_noiso.err; echo "bench noiso rc=$?"
tail -n 3 gpurun_out/t_gpu_tests.log
