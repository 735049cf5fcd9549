# Records of the resident kernel round on a 1-GPU box:   gpurun -- bash tools/resident_round.sh
set -u
mkdir -p gpurun_out
timeout 500 python -m pytest tests -q -m gpu > gpurun_out/r02_gpu_tests_final.txt 2>&1; tail -4 gpurun_out/r02_gpu_tests_final.txt
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 200 python tools/bench_configs.py > gpurun_out/r02_other_configs.jsonl 2> gpurun_out/r02_other_configs.err; grep -c . gpurun_out/r02_other_configs.jsonl; tail -2 gpurun_out/r02_other_configs.err
timeout 100 python tools/bench_resident.py > gpurun_out/r02_resident_bench.jsonl 2> gpurun_out/r02_resident_bench.err; grep -c . gpurun_out/r02_resident_bench.jsonl
timeout 200 python tools/bnn_trajectory_drift.py --variants resident > gpurun_out/r02_bnn_trajectory_drift_resident.jsonl 2>&1; cut -c1-400 gpurun_out/r02_bnn_trajectory_drift_resident.jsonl
timeout 300 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err; cut -c1-200 gpurun_out/r02_bench_n1.json; tail -2 gpurun_out/r02_bench_n1.err
timeout 150 ncu --set full --clock-control none --import-source on -k regex:bnn_sghmc_resident -s 1 -c 1 -f -o gpurun_out/r02_resident python tools/bench_resident.py --chains 148 --overlap 1 --steps 32 --reps 1 > gpurun_out/r02_resident_ncu.out 2>&1; ls -la gpurun_out/r02_resident.ncu-rep
