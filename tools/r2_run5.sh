set -x
mkdir -p gpurun_out
KB=tools/kbench/_build/kbench; CB=tools/kbench/_build/consts.bin
RBM_GRAM_VARIANT=4 timeout 60 $KB $CB gram32 12500000 5 > gpurun_out/plain_tc.log 2>&1 && \
RBM_GRAM_VARIANT=4 timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_regressor_gram_tc -s 4 -c 1 -o gpurun_out/r2f_gram32_tc $KB $CB gram32 12500000 5 > gpurun_out/ncu_tc.log 2>&1
tail -3 gpurun_out/ncu_tc.log
