mkdir -p gpurun_out
echo "== full"
timeout 150 python tools/variant_bench.py --variants 2 --n 4096,32768 --reps 2 2>&1 | tee gpurun_out/roles2_full.log
for exp in no_scalar no_cov; do
  echo "== $exp"
  ESKF_B200_LIB=$PWD/dvi_ekf_b200/libeskf_b200_eskf_exp_$exp.so timeout 150 python tools/variant_bench.py --variants 2 --n 4096,32768 --reps 2 2>&1 | tee gpurun_out/roles2_$exp.log
done
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -k "kernel2" 2>&1 | tail -3
