# first GPU contact of a kernel change: quick A/B timing, then the parity tests of the default kernel
set -x
mkdir -p gpurun_out
V=${1:-3}
timeout 120 python tools/variant_bench.py --n 4096 --variants $V --reps 2 > gpurun_out/variant_bench_a.log 2>&1; echo "bench rc=$?"
cat gpurun_out/variant_bench_a.log
timeout 500 python -m pytest tests/test_gpu_parity.py -x -q -k "kernel$V" > gpurun_out/pytest_v$V.log 2>&1; echo "pytest rc=$?"
tail -25 gpurun_out/pytest_v$V.log
timeout 200 python tools/variant_bench.py --variants $V,1 > gpurun_out/variant_bench.log 2>&1; echo "bench rc=$?"
cat gpurun_out/variant_bench.log
