set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_multi.py -q -x 2>&1 | tail -3
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2d_bench_2gpu.log 2>&1; tail -1 gpurun_out/r2d_bench_2gpu.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tools/bench_pcie.py > gpurun_out/r2d_pcie_2gpu.jsonl 2>&1; tail -1 gpurun_out/r2d_pcie_2gpu.jsonl
nvidia-smi topo -m > gpurun_out/r2d_topo_2gpu.txt 2>&1; lscpu | head -25 >> gpurun_out/r2d_topo_2gpu.txt; ls /sys/devices/system/node/ >> gpurun_out/r2d_topo_2gpu.txt
