# Final records on a 1-GPU box:   gpurun -- bash tools/final_round.sh
set -u
mkdir -p gpurun_out
timeout 500 python -m pytest tests -q -m gpu > gpurun_out/r02_gpu_tests_final.txt 2>&1; tail -3 gpurun_out/r02_gpu_tests_final.txt
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 300 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err; cut -c1-200 gpurun_out/r02_bench_n1.json; tail -2 gpurun_out/r02_bench_n1.err
timeout 200 python tools/bench_configs.py > gpurun_out/r02_other_configs.jsonl 2> gpurun_out/r02_other_configs.err; grep -c . gpurun_out/r02_other_configs.jsonl; tail -2 gpurun_out/r02_other_configs.err
timeout 200 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_bnn_resident_gpu.py -q -x -k "iter_host and blocks and 3-8" > gpurun_out/r02_compute_sanitizer_resident_iter_host.txt 2>&1; echo "memcheck rc=$?"; tail -3 gpurun_out/r02_compute_sanitizer_resident_iter_host.txt
