mkdir -p gpurun_out
KB=tools/kbench/_build/kbench; CB=tools/kbench/_build/consts.bin
for v in 0 1 2; do echo "== RBM_LIN_VARIANT=$v"; RBM_LIN_VARIANT=$v timeout 120 $KB $CB lin 1048576 100 2>&1 | tail -2; done > gpurun_out/r2l_kbench_lin.log 2>&1
cat gpurun_out/r2l_kbench_lin.log
for v in 1 2; do RBM_LIN_VARIANT=$v timeout 600 python -m pytest tests/test_gpu_linearize.py tests/test_gpu_replay.py -q -x 2>&1 | tail -3; done
