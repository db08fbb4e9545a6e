#!/bin/bash
# ncu evidence for the five stages of one LM iteration on the bench workload (one B200; never a bench value):
#   1. launch list (gpu__time_duration.sum per launch; cold-cache, serialised: compare SHARES)
#   2. --set full captures of the stage kernels, with source correlation (-lineinfo builds)
TAG=${1:-r2}
CASE=${2:-venice-1778}
PREC=${3:-f64-f64}
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --workload $CASE --precision $PREC --steps 8 --warmup 3 --device-only > gpurun_out/${TAG}_launches.log 2>&1
python scripts/launch_summary.py gpurun_out/${TAG}_launches.csv > gpurun_out/${TAG}_launches.txt 2>&1
ncu --set full --clock-control none --import-source on \
    -k regex:"k_linearize|k_prepare_cams|k_backsubst_points|k_backsubst_tiles|k_cost_tiles|k_pcg_solve|k_point_prepare|k_cam_reduce_prepare" \
    -c 14 -f -o gpurun_out/${TAG}_stages python scripts/prof_one.py $CASE $PREC 6 > gpurun_out/${TAG}_stages.log 2>&1
ls -la gpurun_out/${TAG}_stages.ncu-rep
