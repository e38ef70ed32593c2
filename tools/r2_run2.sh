set -x
mkdir -p gpurun_out
KB=tools/kbench/_build/kbench; CB=tools/kbench/_build/consts.bin
python tools/bench_pcie.py > gpurun_out/r2b_pcie_1gpu.jsonl 2>&1; cat gpurun_out/r2b_pcie_1gpu.jsonl
python tools/bench_e2e_chunk.py > gpurun_out/r2b_e2e_chunk.jsonl 2>&1; cat gpurun_out/r2b_e2e_chunk.jsonl
for v in 0 1; do
  for w in gram gram32; do
    RBM_GRAM_VARIANT=$v $KB $CB $w 12500000 5 > gpurun_out/plain_$w$v.log 2>&1 && \
    RBM_GRAM_VARIANT=$v ncu --set full --clock-control none --import-source on -k regex:k_regressor_gram_ -s 8 -c 1 -o gpurun_out/r2b_${w}_v$v $KB $CB $w 12500000 5 > gpurun_out/ncu_$w$v.log 2>&1
    tail -2 gpurun_out/ncu_$w$v.log
  done
done
ls -la gpurun_out
