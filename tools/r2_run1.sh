set -x
mkdir -p gpurun_out
KB=tools/kbench/_build/kbench; CB=tools/kbench/_build/consts.bin
for v in 0 1; do RBM_GRAM_VARIANT=$v $KB $CB gram 12500000 50 2>&1 | grep -v "^  n=" ; done > gpurun_out/r2a_kbench_gram.log 2>&1
for v in 0 1; do RBM_GRAM_VARIANT=$v $KB $CB gram32 12500000 50 2>&1 ; done >> gpurun_out/r2a_kbench_gram.log 2>&1
cat gpurun_out/r2a_kbench_gram.log
timeout 900 python -m pytest tests -q -m gpu -x 2>&1 | tail -8
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2a_bench.log 2>&1; tail -1 gpurun_out/r2a_bench.log
