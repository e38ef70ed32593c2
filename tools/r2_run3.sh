set -x
mkdir -p gpurun_out
KB=tools/kbench/_build/kbench; CB=tools/kbench/_build/consts.bin
: > gpurun_out/r2c_kbench_gram.log
for v in 0 1 2 3; do echo "== RBM_GRAM_VARIANT=$v" >> gpurun_out/r2c_kbench_gram.log; RBM_GRAM_VARIANT=$v $KB $CB gram 12500000 50 >> gpurun_out/r2c_kbench_gram.log 2>&1; done
for v in 0 1 2 3; do echo "== RBM_GRAM_VARIANT=$v" >> gpurun_out/r2c_kbench_gram.log; RBM_GRAM_VARIANT=$v $KB $CB gram32 12500000 50 >> gpurun_out/r2c_kbench_gram.log 2>&1; done
cat gpurun_out/r2c_kbench_gram.log
RBM_GRAM_VARIANT=3 timeout 600 python -m pytest tests/test_gpu_regressor.py -q -x 2>&1 | tail -4
RBM_GRAM_VARIANT=2 timeout 600 python -m pytest tests/test_gpu_regressor.py tests/test_gpu_host_entries.py tests/test_gpu_linearize.py tests/test_gpu_bench_contract.py -q -x 2>&1 | tail -4
