# compute-sanitizer memcheck of the layer kernels (FFMA + tcgen05) and the round-2 ncu captures:
#   gpurun -- bash tools/sanitize_and_profile_r02.sh
set -u
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_mlp_gpu.py -q -m gpu -k "tensor_core or predict or generic" -x 2>&1 | tail -15 > gpurun_out/r02_sanitizer_mlp.txt
echo "exit code: $?" >> gpurun_out/r02_sanitizer_mlp.txt
tail -8 gpurun_out/r02_sanitizer_mlp.txt
bash tools/ncu_capture_r02.sh > gpurun_out/r02_ncu_capture.log 2>&1; tail -4 gpurun_out/r02_ncu_capture.log
