#!/bin/bash
# round 2, GPU call AC: full GPU suite + default bench with the final build; memcheck / racecheck over the column-stencil transport test
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/ac_gpu_tests.log 2>&1; echo "gpu tests rc=$?"; tail -n 3 gpurun_out/ac_gpu_tests.log
timeout 900 python bench.py > gpurun_out/ac_bench.json 2> gpurun_out/ac_bench.err; echo "bench rc=$?"
timeout 900 python bench.py --impl reference > gpurun_out/ac_bench_reference.json 2> gpurun_out/ac_bench_reference.err; echo "reference arm rc=$?"
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "column_stencils and False" > gpurun_out/ac_san_memcheck.log 2>&1; echo "memcheck rc=$?"; grep "ERROR SUMMARY\|passed\|failed" gpurun_out/ac_san_memcheck.log | tail -2
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 7 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "column_stencils and block and False" > gpurun_out/ac_san_racecheck.log 2>&1; echo "racecheck rc=$?"; grep "RACECHECK SUMMARY\|passed\|failed" gpurun_out/ac_san_racecheck.log | tail -2
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/ac_smoke.log 2>&1; echo "smoke rc=$?"; tail -n 1 gpurun_out/ac_smoke.log
