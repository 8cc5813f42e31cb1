#!/bin/bash
# round 2, GPU call Y: evidence for the final kernels -- ncu --set full (config 4 and config 2), per-line profiles, the launch list of
# the bench command, compute-sanitizer (memcheck, racecheck, initcheck) over the tile-path parity tests
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_sym_tile|k_gather_tile|k_adj_place|k_h8_diffusion' -s 4 -c 4 -f -o gpurun_out/r02_c4_final2 python profiles/prof_diffusion.py 256 > gpurun_out/y_ncu_c4.log 2>&1; echo "ncu c4 rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_sym_tile|k_gather_tile|k_adj_place|k_h8_elastic' -s 4 -c 4 -f -o gpurun_out/r02_c2_final2 python profiles/prof_elastic.py 128 > gpurun_out/y_ncu_c2.log 2>&1; echo "ncu c2 rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_bench_final.csv python bench.py --steps 2 --warmup 3 > gpurun_out/y_bench_under_ncu.log 2>&1; echo "launch list rc=$?"
for c in c4 c2; do python profiles/ncu_summary.py gpurun_out/r02_${c}_final2.ncu-rep > gpurun_out/y_ncu_${c}_summary.txt 2>&1; done
python profiles/lineprof.py gpurun_out/r02_c4_final2.ncu-rep k_sym_tile 40 > gpurun_out/y_lines_sym_tile.txt 2>&1
python profiles/lineprof.py gpurun_out/r02_c4_final2.ncu-rep k_gather_tile 30 > gpurun_out/y_lines_gather_c4.txt 2>&1
python profiles/lineprof.py gpurun_out/r02_c2_final2.ncu-rep k_gather_tile 30 > gpurun_out/y_lines_gather_c2.txt 2>&1
python profiles/lineprof.py gpurun_out/r02_c2_final2.ncu-rep k_h8_elastic 30 > gpurun_out/y_lines_elastic_iso.txt 2>&1
python profiles/launch_summary.py gpurun_out/r02_launches_bench_final.csv > gpurun_out/y_launch_summary.txt 2>&1
rm -f gpurun_out/r02_c2_final2.ncu-rep
sz=$(du -sm gpurun_out | cut -f1); if [ "$sz" -gt 55 ]; then rm -f gpurun_out/r02_c4_final2.ncu-rep; echo "dropped the c4 report (size)"; fi
SEL="tests/test_gpu_tile.py"
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest $SEL -m gpu -q -x -k "not at_size" > gpurun_out/y_san_memcheck.log 2>&1; echo "memcheck rc=$?"
timeout 1500 compute-sanitizer --tool racecheck --error-exitcode 7 python -m pytest $SEL -m gpu -q -x -k "parity or partition or many_tiles" > gpurun_out/y_san_racecheck.log 2>&1; echo "racecheck rc=$?"
timeout 1500 compute-sanitizer --tool initcheck --error-exitcode 7 python -m pytest $SEL -m gpu -q -x -k "parity or partition" > gpurun_out/y_san_initcheck.log 2>&1; echo "initcheck rc=$?"
for t in memcheck racecheck initcheck; do echo "== $t"; grep -c "=========" gpurun_out/y_san_$t.log; tail -n 3 gpurun_out/y_san_$t.log; done
