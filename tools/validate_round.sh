# Round validation on a 1-GPU box:   gpurun -- bash tools/validate_round.sh
set -u
mkdir -p gpurun_out
python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err
cut -c1-400 gpurun_out/r02_bench_n1.json; tail -3 gpurun_out/r02_bench_n1.err
python tools/bench_k4.py --variants 0,10,11,13,14,15,16 > gpurun_out/r02_k4_variants.jsonl 2>&1; cat gpurun_out/r02_k4_variants.jsonl
python tools/bench_mlp.py --chains 1,8,64,256 > gpurun_out/r02_mlp_wide.jsonl 2> gpurun_out/r02_mlp_wide.err; cut -c1-260 gpurun_out/r02_mlp_wide.jsonl
python tools/bench_configs.py > gpurun_out/r02_other_configs.jsonl 2> gpurun_out/r02_other_configs.err; tail -3 gpurun_out/r02_other_configs.jsonl | cut -c1-300
python -m pytest tests -q -m gpu 2>&1 | tail -6 > gpurun_out/r02_gpu_tests.txt; cat gpurun_out/r02_gpu_tests.txt
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
