mkdir -p gpurun_out
KB=tools/kbench/_build/kbench; CB=tools/kbench/_build/consts.bin
for d in 8 1; do
RBM_TC_DEBUG=$d RBM_GRAM_VARIANT=4 timeout 60 $KB $CB gram32 12500000 5 > gpurun_out/plain_tc$d.log 2>&1 && \
RBM_TC_DEBUG=$d RBM_GRAM_VARIANT=4 timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_regressor_gram_tc -s 4 -c 1 -o gpurun_out/r2i_gram32_tc_dbg$d $KB $CB gram32 12500000 5 > gpurun_out/ncu_tc$d.log 2>&1
tail -1 gpurun_out/ncu_tc$d.log
done
