# Last GPU pass of a round (1 GPU):   gpurun -- bash tools/final_checks.sh
#   memcheck of the newest kernels, the ncu captures that stamp profiles/ncu_traffic.json, the validation run
set -u
mkdir -p gpurun_out
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_mlp_gpu.py tests/test_bnn_gpu.py -q -m gpu -k "tensor_core or predict or generic or fused or k4_matches or golden" -x 2>&1 | tail -6 > gpurun_out/r02_sanitizer.txt
echo "exit code: $?" >> gpurun_out/r02_sanitizer.txt; tail -4 gpurun_out/r02_sanitizer.txt
timeout 1200 bash tools/ncu_capture_r02.sh > gpurun_out/r02_ncu_capture.log 2>&1; tail -3 gpurun_out/r02_ncu_capture.log
bash tools/validate_round.sh
