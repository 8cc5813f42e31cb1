#!/bin/bash
# round 2, GPU call AJ: last validation of the final build: full GPU suite + smoke
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/aj_gpu_tests.log 2>&1; echo "gpu tests rc=$?"; tail -n 2 gpurun_out/aj_gpu_tests.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/aj_smoke.log 2>&1; echo "smoke rc=$?"; tail -n 1 gpurun_out/aj_smoke.log
