# Records of the resident-kernel work on a 1-GPU box:   gpurun -- bash tools/resident_round.sh
set -u
mkdir -p gpurun_out
timeout 500 python -m pytest tests -q -m gpu > gpurun_out/r02_gpu_tests_final.txt 2>&1; tail -4 gpurun_out/r02_gpu_tests_final.txt
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 300 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err; cut -c1-200 gpurun_out/r02_bench_n1.json; tail -2 gpurun_out/r02_bench_n1.err
timeout 200 python tools/bench_configs.py > gpurun_out/r02_other_configs.jsonl 2> gpurun_out/r02_other_configs.err; grep -c . gpurun_out/r02_other_configs.jsonl; tail -2 gpurun_out/r02_other_configs.err
timeout 100 python tools/bench_resident.py > gpurun_out/r02_resident_bench.jsonl 2> gpurun_out/r02_resident_bench.err; grep -c . gpurun_out/r02_resident_bench.jsonl
timeout 240 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_bnn_resident_gpu.py -q -x -k "gradient or block or noise or run_equals or iter_host" > gpurun_out/r02_compute_sanitizer_resident.txt 2>&1; echo "memcheck rc=$?"; tail -4 gpurun_out/r02_compute_sanitizer_resident.txt
timeout 200 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_bnn_resident_gpu.py -q -x -k "block_of_steps and overlap-True" > gpurun_out/r02_racecheck_resident.txt 2>&1; echo "racecheck rc=$?"; tail -4 gpurun_out/r02_racecheck_resident.txt
