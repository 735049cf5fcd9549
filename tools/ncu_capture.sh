#!/bin/bash
# Profiling pass on the GPU box (1 GPU): launch list + full captures of K4 and K1.
#   gpurun -- bash tools/ncu_capture.sh <tag>
# Numbers printed by bench.py under ncu are NOT bench values.
set -u
TAG=${1:-r01}
mkdir -p gpurun_out
CMD="python bench.py --steps 20 --warmup 3 --no-cpu-baseline --chains-per-gpu 8192"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file gpurun_out/${TAG}_launches.csv $CMD > gpurun_out/${TAG}_launches.out 2>&1
ncu --set full --clock-control none --import-source on -k regex:bnn_mma_kernel -s 4 -c 2 \
    -f -o gpurun_out/${TAG}_k4 $CMD > gpurun_out/${TAG}_k4.out 2>&1
ncu --set full --clock-control none --import-source on -k regex:sghmc_update_kernel -s 4 -c 2 \
    -f -o gpurun_out/${TAG}_k1 $CMD > gpurun_out/${TAG}_k1.out 2>&1
ls -la gpurun_out/
