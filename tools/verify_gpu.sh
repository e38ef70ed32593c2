# Round-end style verification on one GPU box:  gpurun --timeout 1800 -- 'bash tools/verify_gpu.sh'
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu -x 2>&1 | tail -4
timeout 300 python -c "import __graft_entry__ as g; g.build(); g.smoke(); print('SMOKE OK')" 2>&1 | tail -2 | cut -c1-300
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/verify_bench_default.log 2>&1; tail -1 gpurun_out/verify_bench_default.log | cut -c1-300
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/verify_bench_ref.log 2>&1; tail -1 gpurun_out/verify_bench_ref.log | cut -c1-200
