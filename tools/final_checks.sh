set -u
mkdir -p gpurun_out
python -m pytest tests -q -m gpu -x 2>&1 | tail -4 > gpurun_out/r02_gpu_tests.txt; cat gpurun_out/r02_gpu_tests.txt
timeout 1200 bash tools/ncu_capture_r02.sh > gpurun_out/r02_ncu_capture.log 2>&1; tail -3 gpurun_out/r02_ncu_capture.log
python tools/bench_k4.py --variants 0,10,11,13,14,15,16 > gpurun_out/r02_k4_variants.jsonl 2>&1; cut -c1-70 gpurun_out/r02_k4_variants.jsonl
