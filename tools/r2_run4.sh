set -x
mkdir -p gpurun_out
KB=tools/kbench/_build/kbench; CB=tools/kbench/_build/consts.bin
RBM_GRAM_VARIANT=4 timeout 60 $KB $CB gram32 1000000 5 > gpurun_out/r2e_tc_small.log 2>&1; echo "rc=$?" >> gpurun_out/r2e_tc_small.log; cat gpurun_out/r2e_tc_small.log
RBM_GRAM_VARIANT=4 timeout 60 $KB $CB gram32 12500000 50 > gpurun_out/r2e_tc.log 2>&1; echo "rc=$?" >> gpurun_out/r2e_tc.log; cat gpurun_out/r2e_tc.log
RBM_GRAM_VARIANT=2 timeout 60 $KB $CB gram32 12500000 50 2>&1 | tail -3
RBM_GRAM_VARIANT=4 timeout 600 python -m pytest tests/test_gpu_regressor.py -q -x -k "fp32 or identification or config1" 2>&1 | tail -5
