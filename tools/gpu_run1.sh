set -x
mkdir -p gpurun_out
timeout 120 python tools/variant_bench.py --n 4096 --variants 2 --reps 2 > gpurun_out/variant_bench_a.log 2>&1; echo "bench v2 rc=$?"
cat gpurun_out/variant_bench_a.log
timeout 400 python -m pytest tests/test_gpu_parity.py -x -q -k "kernel2" > gpurun_out/pytest_v2.log 2>&1; echo "pytest v2 rc=$?"
tail -25 gpurun_out/pytest_v2.log
timeout 200 python tools/variant_bench.py --variants 2 > gpurun_out/variant_bench.log 2>&1; echo "bench rc=$?"
cat gpurun_out/variant_bench.log
