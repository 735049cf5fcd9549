#!/bin/bash
# Round-2 profiling pass on the GPU box (1 GPU): launch list of the bench + full captures of K4 (default
# accuracy mode), K1 and the tcgen05 layer kernel of the wide net.   gpurun -- bash tools/ncu_capture_r02.sh
# Numbers printed by bench.py under ncu are NOT bench values.
set -u
mkdir -p gpurun_out
CMD="python bench.py --steps 20 --warmup 3 --no-cpu-baseline --chains-per-gpu 8192"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file gpurun_out/r02_launches.csv $CMD > gpurun_out/r02_launches.out 2>&1
ncu --set full --clock-control none --import-source on -k regex:bnn_mma_kernel -s 4 -c 2 \
    -f -o gpurun_out/r02_k4 $CMD > gpurun_out/r02_k4.out 2>&1
ncu --set full --clock-control none --import-source on -k regex:sghmc_update_kernel -s 4 -c 2 \
    -f -o gpurun_out/r02_k1 $CMD > gpurun_out/r02_k1.out 2>&1
ncu --set full --clock-control none --import-source on -k regex:mlp_gemm_umma_kernel -s 2 -c 4 \
    -f -o gpurun_out/r02_mlp_umma python tools/bench_mlp.py --chains 256 --iters 2 --layers tcgen05 > gpurun_out/r02_mlp_umma.out 2>&1
ls -la gpurun_out/*.ncu-rep
