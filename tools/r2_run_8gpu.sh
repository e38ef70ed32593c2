set -x
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 900 $TR --nproc-per-node 8 --master-port 29521 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r2k_bench_8gpu.log 2>&1; tail -1 gpurun_out/r2k_bench_8gpu.log | cut -c1-600
timeout 600 $TR --nproc-per-node 8 --master-port 29522 tools/sweep_rnea.py --out gpurun_out/r2k_sweep_rnea_8gpu.jsonl > gpurun_out/r2k_sweep8.log 2>&1; tail -3 gpurun_out/r2k_sweep8.log | cut -c1-300
timeout 600 $TR --nproc-per-node 4 --master-port 29523 tools/sweep_rnea.py --dtypes f32 --out gpurun_out/r2k_sweep_rnea_4gpu.jsonl > gpurun_out/r2k_sweep4.log 2>&1; tail -1 gpurun_out/r2k_sweep4.log | cut -c1-300
timeout 600 $TR --nproc-per-node 2 --master-port 29524 tools/sweep_rnea.py --dtypes f32 --out gpurun_out/r2k_sweep_rnea_2gpu.jsonl > gpurun_out/r2k_sweep2.log 2>&1; tail -1 gpurun_out/r2k_sweep2.log | cut -c1-300
timeout 300 $TR --nproc-per-node 8 --master-port 29525 tools/bench_pcie.py > gpurun_out/r2k_pcie_8gpu.jsonl 2>&1; tail -1 gpurun_out/r2k_pcie_8gpu.jsonl
timeout 300 $TR --nproc-per-node 4 --master-port 29526 tools/bench_pcie.py > gpurun_out/r2k_pcie_4gpu.jsonl 2>&1; tail -1 gpurun_out/r2k_pcie_4gpu.jsonl
nvidia-smi topo -m > gpurun_out/r2k_topo_8gpu.txt 2>&1; lscpu | grep -i "socket\|numa\|^CPU(s)\|model name" >> gpurun_out/r2k_topo_8gpu.txt; free -g >> gpurun_out/r2k_topo_8gpu.txt
