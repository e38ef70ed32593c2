set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests -q -m gpu -x 2>&1 | tail -4
timeout 300 python -c "import __graft_entry__ as g; g.build(); g.smoke(); print('SMOKE OK')" 2>&1 | tail -2
timeout 300 python bench.py > gpurun_out/r1d_bench_default.log 2>&1; tail -1 gpurun_out/r1d_bench_default.log | cut -c1-400
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r1d_bench_ref.log 2>&1; tail -1 gpurun_out/r1d_bench_ref.log | cut -c1-200
timeout 200 python bench.py --workload linearize --steps 20 --warmup 5 --no-cpu > gpurun_out/r1d_bench_lin.log 2>&1; tail -1 gpurun_out/r1d_bench_lin.log | cut -c1-300
timeout 200 python bench.py --workload gram --steps 20 --warmup 5 --no-cpu > gpurun_out/r1d_bench_gram.log 2>&1; tail -1 gpurun_out/r1d_bench_gram.log | cut -c1-300
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_linearize -c 2 -f -o gpurun_out/prof_r1d_linearize python bench.py --workload linearize --steps 2 --warmup 3 --no-cpu > gpurun_out/r1d_ncu_lin.log 2>&1; tail -2 gpurun_out/r1d_ncu_lin.log | cut -c1-200
