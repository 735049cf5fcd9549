set -u
mkdir -p gpurun_out
python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err
cut -c1-600 gpurun_out/r02_bench_n1.json; tail -3 gpurun_out/r02_bench_n1.err
python tools/bench_rsghmc_sweep.py > gpurun_out/r02_sweep_n1.jsonl 2> gpurun_out/r02_sweep_n1.err; tail -2 gpurun_out/r02_sweep_n1.jsonl | cut -c1-300
python tools/bench_configs.py > gpurun_out/r02_other_configs.jsonl 2> gpurun_out/r02_other_configs.err; tail -3 gpurun_out/r02_other_configs.jsonl | cut -c1-300
python -m pytest tests -q -m gpu 2>&1 | tail -6 > gpurun_out/r02_gpu_tests.txt; cat gpurun_out/r02_gpu_tests.txt
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
