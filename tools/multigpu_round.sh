# Multi-GPU records of a round on an N-GPU box:   gpurun --gpus N -- bash tools/multigpu_round.sh N
#   * the NCCL test of the one collective of the path (tests/test_diagnostics_multigpu.py, needs >= 2 GPUs)
#   * BASELINE.json configs[4]: relativistic update sweep D = 1e5 ... 1e9, weak and strong, under torchrun
#   * bench.py --gpus N under torchrun (chain-steps/s, e2e, R-hat / ESS of all chains with the all-reduce)
set -u
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi -L | head -8 > gpurun_out/r02_mg${N}_gpus.txt
python -m pytest tests/test_diagnostics_multigpu.py -q -m gpu -v 2>&1 | tail -6 > gpurun_out/r02_mg${N}_diag_test.txt
cat gpurun_out/r02_mg${N}_diag_test.txt
python -m torch.distributed.run --nnodes=1 --nproc-per-node ${N} --master-addr 127.0.0.1 --master-port 29511 tools/bench_rsghmc_sweep.py > gpurun_out/r02_sweep_n${N}.jsonl 2> gpurun_out/r02_sweep_n${N}.err
tail -3 gpurun_out/r02_sweep_n${N}.jsonl | cut -c1-400
python -m torch.distributed.run --nnodes=1 --nproc-per-node ${N} --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus ${N} --steps 20 --warmup 5 > gpurun_out/r02_bench_n${N}.json 2> gpurun_out/r02_bench_n${N}.err
cut -c1-1500 gpurun_out/r02_bench_n${N}.json; tail -3 gpurun_out/r02_bench_n${N}.err
