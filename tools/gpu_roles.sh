# role split of the default kernel: full / scalar roles only / covariance role only (experiment builds)
mkdir -p gpurun_out
V=${1:-3}
echo "== full"
timeout 150 python tools/variant_bench.py --variants $V --n 4096 --reps 2 2>&1 | tee gpurun_out/roles${V}_full.log
for exp in no_scalar no_cov; do
  echo "== $exp"
  ESKF_B200_LIB=$PWD/dvi_ekf_b200/libeskf_b200_eskf_exp_$exp.so timeout 150 python tools/variant_bench.py --variants $V --n 4096 --reps 2 2>&1 | tee gpurun_out/roles${V}_$exp.log
done
