mkdir -p gpurun_out
KB=tools/kbench/_build/kbench; CB=tools/kbench/_build/consts.bin
BENCH="python bench.py --steps 20 --warmup 5 --no-extra --no-cpu"
$BENCH > gpurun_out/plain_bench.log 2>&1 && \
ncu --set full --cache-control none --clock-control none --import-source on -k regex:k_rnea_fast_soa -s 40 -c 3 -o gpurun_out/r2m_rnea_f64_steady $BENCH > gpurun_out/ncu_rnea.log 2>&1
tail -1 gpurun_out/ncu_rnea.log
$BENCH > gpurun_out/plain_bench2.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2m_launches_bench_f64.csv $BENCH > gpurun_out/ncu_launches.log 2>&1
for w in gram gram32; do
  timeout 60 $KB $CB $w 12500000 5 > gpurun_out/plain_$w.log 2>&1 && \
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_regressor_gram_pipe -s 8 -c 1 -o gpurun_out/r2m_${w}_v2 $KB $CB $w 12500000 5 > gpurun_out/ncu_$w.log 2>&1
  tail -1 gpurun_out/ncu_$w.log
done
RBM_LIN_VARIANT=1 timeout 60 $KB $CB lin 1048576 5 > gpurun_out/plain_lin.log 2>&1 && \
RBM_LIN_VARIANT=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_linearize_pair -s 4 -c 1 -o gpurun_out/r2m_lin_pair $KB $CB lin 1048576 5 > gpurun_out/ncu_lin.log 2>&1
tail -1 gpurun_out/ncu_lin.log
RBM_LIN_VARIANT=0 timeout 60 $KB $CB lin 1048576 5 > gpurun_out/plain_lin0.log 2>&1 && \
RBM_LIN_VARIANT=0 timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_linearize_fast -s 4 -c 1 -o gpurun_out/r2m_lin_single $KB $CB lin 1048576 5 > gpurun_out/ncu_lin0.log 2>&1
tail -1 gpurun_out/ncu_lin0.log
