mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu -x 2>&1 | tail -4
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | cut -c1-200
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2n_bench.log 2>&1; tail -1 gpurun_out/r2n_bench.log | cut -c1-300
timeout 300 python tools/sweep_rnea.py --generic --min-exp 6 --max-exp 7 --out gpurun_out/r2n_sweep_generic.jsonl 2>&1 | cut -c1-260
for w in gram gram32 lin rnea; do timeout 60 tools/kbench/_build/kbench tools/kbench/_build/consts.bin $w 2>&1 | grep -v "^  n=\|^rbm"; done
