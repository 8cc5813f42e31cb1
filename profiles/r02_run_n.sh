#!/bin/bash
# round 2, GPU call N: persistent k_h8_elastic with bulk (copy-engine) stores; co-residency knob for k_h8_diffusion beside k_sym_tile;
# full ncu captures of the final kernels (config 4 and config 2); launch list of the bench command
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
# (first attempt of this call also ran tests/test_gpu_parity.py + tests/test_gpu_tile.py: 155 passed; its outputs exceeded the 64 MiB
# the box copies back, so this version summarises the reports on the box and keeps only what fits)
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/n_bench.json 2> gpurun_out/n_bench.err; echo "bench rc=$?"
FEGPU_ELASTIC_BULK=0 timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/n_bench_nobulk.json 2> gpurun_out/n_bench_nobulk.err; echo "bench nobulk rc=$?"
FEGPU_DIFF_PAD_KB=120 timeout 600 python bench.py --steps 10 --warmup 3 --no-secondary > gpurun_out/n_bench_pad120.json 2> gpurun_out/n_bench_pad120.err; echo "bench pad rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_sym_tile|k_gather_tile|k_adj_place|k_h8_diffusion' -s 4 -c 4 -f -o gpurun_out/r02_c4_final python profiles/prof_diffusion.py 256 > gpurun_out/n_ncu_c4.log 2>&1; echo "ncu c4 rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_gather_tile|k_h8_elastic' -s 2 -c 2 -f -o gpurun_out/r02_c2_final python profiles/prof_elastic.py 128 > gpurun_out/n_ncu_c2.log 2>&1; echo "ncu c2 rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_bench.csv python bench.py --steps 2 --warmup 3 > gpurun_out/n_bench_under_ncu.log 2>&1; echo "launch list rc=$?"
for c in c4 c2; do
  python profiles/ncu_summary.py gpurun_out/r02_${c}_final.ncu-rep > gpurun_out/n_ncu_${c}_summary.txt 2>&1
done
python profiles/lineprof.py gpurun_out/r02_c4_final.ncu-rep k_sym_tile 60 > gpurun_out/n_lines_sym_tile.txt 2>&1
python profiles/lineprof.py gpurun_out/r02_c2_final.ncu-rep k_gather_tile 40 > gpurun_out/n_lines_gather_c2.txt 2>&1
python profiles/lineprof.py gpurun_out/r02_c2_final.ncu-rep k_h8_elastic 40 > gpurun_out/n_lines_elastic.txt 2>&1
python profiles/lineprof.py gpurun_out/r02_c4_final.ncu-rep k_gather_tile 40 > gpurun_out/n_lines_gather_c4.txt 2>&1
sz=$(du -sm gpurun_out | cut -f1); echo "gpurun_out: ${sz} MiB"
if [ "$sz" -gt 58 ]; then rm -f gpurun_out/r02_c2_final.ncu-rep; echo "dropped the c2 report (size)"; fi
sz=$(du -sm gpurun_out | cut -f1)
if [ "$sz" -gt 58 ]; then rm -f gpurun_out/r02_c4_final.ncu-rep; echo "dropped the c4 report (size)"; fi
