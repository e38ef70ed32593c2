mkdir -p gpurun_out
KB=tools/kbench/_build/kbench; CB=tools/kbench/_build/consts.bin
for d in 0 1 2 4 6 8 10 14; do echo "== RBM_TC_DEBUG=$d"; RBM_TC_DEBUG=$d RBM_GRAM_VARIANT=4 timeout 60 $KB $CB gram32 12500000 30 2>&1 | grep "TMA pipeline"; done > gpurun_out/r2h_tc_debug.log 2>&1
cat gpurun_out/r2h_tc_debug.log
