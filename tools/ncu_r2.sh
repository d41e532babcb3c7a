#!/bin/bash
# ncu --set full captures of the hot kernels at the bench size (one GPU).
# Run under gpurun; the plain run of the same command comes first.
# The reports are exported to CSV on the box (gpurun_out/ is capped at 64 MiB)
# and only the small fused-smoother report travels back.
set -x
mkdir -p gpurun_out
CMD="python tools/microbench.py --only ncu --reps 1"
$CMD > gpurun_out/ncu_plain.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on \
    -k regex:'k_gs_fused' -s 6 -c 2 \
    -o gpurun_out/r2_fused -f $CMD > gpurun_out/ncu_fused.log 2>&1
echo ncu fused rc=$?
ncu -i gpurun_out/r2_fused.ncu-rep --page raw --csv > gpurun_out/r2_fused_raw.csv 2>/dev/null
ncu -i gpurun_out/r2_fused.ncu-rep --page source --csv > gpurun_out/r2_fused_source.csv 2>/dev/null
ncu -i gpurun_out/r2_fused.ncu-rep --page source --csv --print-source sass > gpurun_out/r2_fused_sass.csv 2>/dev/null
ncu --set full --clock-control none \
    -k regex:'k_space_spmm|k_time_apply|k_wavelet' -c 24 \
    -o /tmp/r2_rest -f $CMD > gpurun_out/ncu_rest.log 2>&1
echo ncu rest rc=$?
ncu -i /tmp/r2_rest.ncu-rep --page raw --csv > gpurun_out/r2_rest_raw.csv 2>/dev/null
ls -la gpurun_out /tmp/r2_rest.ncu-rep
# keep the report only if it fits comfortably
if [ $(stat -c %s gpurun_out/r2_fused.ncu-rep) -gt 30000000 ]; then rm gpurun_out/r2_fused.ncu-rep; fi
du -sh gpurun_out
