#!/bin/bash
# Round-2 evidence run on one B200 (under gpurun): tests, both bench arms, the
# operator microbenchmark of BASELINE configs[2], a deep-wavefront numbering,
# the launch list and ncu --set full of the hot kernels.
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/r2k_tests.log 2>&1; echo tests rc=$?
python bench.py > gpurun_out/r2k_bench.json 2> gpurun_out/r2k_bench.err; echo bench rc=$?
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2k_bench_reference.json 2> gpurun_out/r2k_bench_reference.err; echo ref rc=$?
python tools/microbench.py > gpurun_out/r2k_micro.log 2>&1
python -m spacetime_fullgrid_parallel_b200.heateq_mpi_timing --J_time 9 --J_space 9 --wavelettransform composite > gpurun_out/r2k_timing_9_9_composite.log 2>&1; echo timing rc=$?
python -m spacetime_fullgrid_parallel_b200.heateq_mpi_timing --J_time 9 --J_space 9 > gpurun_out/r2k_timing_9_9_original.log 2>&1; echo timing rc=$?
python bench.py --J_time 4 --J_space 10 --steps 3 --warmup 1 --no-cpu > gpurun_out/r2k_bench_Js10.json 2> gpurun_out/r2k_bench_Js10.err; echo Js10 rc=$?
timeout 600 python bench.py --order lex --steps 2 --warmup 1 --no-e2e --no-cpu > gpurun_out/r2k_bench_lex.json 2> gpurun_out/r2k_bench_lex.err; echo lex rc=$?
# launch list (same command first without ncu)
python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu > gpurun_out/r2k_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv \
    --log-file gpurun_out/r2k_launches.csv python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu > gpurun_out/r2k_ncu1.log 2>&1
echo launches rc=$?
CMD="python tools/microbench.py --only ncu --reps 1"
$CMD > gpurun_out/r2k_ncu_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'k_gs_fused' -s 4 -c 2 \
    -o gpurun_out/r2k_fused -f $CMD > gpurun_out/r2k_ncu2.log 2>&1
ncu -i gpurun_out/r2k_fused.ncu-rep --page raw --csv > gpurun_out/r2k_fused_raw.csv 2>/dev/null
ncu -i gpurun_out/r2k_fused.ncu-rep --page source --csv > gpurun_out/r2k_fused_source.csv 2>/dev/null
ncu --set full --clock-control none -k regex:'k_space_spmm|k_time_apply|k_wavelet' -c 24 \
    -o /tmp/r2k_rest -f $CMD > gpurun_out/r2k_ncu3.log 2>&1
ncu -i /tmp/r2k_rest.ncu-rep --page raw --csv > gpurun_out/r2k_rest_raw.csv 2>/dev/null
python tools/microbench.py --only fused --reps 1 > gpurun_out/r2k_ncu_plain2.log 2>&1 && \
ncu --set full --clock-control none -k regex:'k_space_spmm_gk|k_mg_coarse' -c 6 \
    -o /tmp/r2k_mg -f python tools/microbench.py --only fused --reps 1 > gpurun_out/r2k_ncu4.log 2>&1
ncu -i /tmp/r2k_mg.ncu-rep --page raw --csv > gpurun_out/r2k_mg_raw.csv 2>/dev/null
if [ $(stat -c %s gpurun_out/r2k_fused.ncu-rep) -gt 20000000 ]; then rm gpurun_out/r2k_fused.ncu-rep; fi
du -sh gpurun_out
