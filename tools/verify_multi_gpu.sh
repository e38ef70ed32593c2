# Multi-GPU verification on one box with N GPUs:  gpurun --gpus 8 --timeout 2400 -- 'bash tools/verify_multi_gpu.sh 8'
# (bench line with extra.gram = configs[2] + NCCL all-reduce, the configs[4] sweep, the host copy ceiling)
N=${1:-2}
set -x
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --nproc-per-node $N"
timeout 900 $TR --master-port 29521 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_${N}gpu.log 2>&1; tail -1 gpurun_out/bench_${N}gpu.log | cut -c1-400
timeout 600 $TR --master-port 29522 tools/sweep_rnea.py --out gpurun_out/sweep_rnea_${N}gpu.jsonl > gpurun_out/sweep_${N}gpu.log 2>&1; tail -2 gpurun_out/sweep_${N}gpu.log | cut -c1-300
timeout 300 $TR --master-port 29525 tools/bench_pcie.py > gpurun_out/pcie_${N}gpu.jsonl 2>&1; tail -1 gpurun_out/pcie_${N}gpu.jsonl
timeout 600 python -m pytest tests/test_gpu_multi.py -q 2>&1 | tail -2
