# Round-end evidence: plain bench, ncu launch list of the same command, ncu --set full of the persistent kernel on the
# bench workload (4096 filters, whole trajectory, noise on).  Outputs in gpurun_out/, summarised under profiles/.
set -x
mkdir -p gpurun_out
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-configs > gpurun_out/ncu_launch.log 2>&1; echo "ncu launches rc=$?"
export ESKF_B200_VARIANT=3
timeout 900 ncu --set full --import-source on --clock-control none -k regex:eskf_kernel -s 1 -c 1 -o gpurun_out/prof_bench -f \
  python tools/profile_run.py --filters 4096 --frames 0 --passes 2 > gpurun_out/ncu_full_bench.log 2>&1; echo "ncu full rc=$?"
ncu -i gpurun_out/prof_bench.ncu-rep --page raw --csv > gpurun_out/raw_bench.csv 2>/dev/null
ncu -i gpurun_out/prof_bench.ncu-rep --page source --csv > gpurun_out/source_bench.csv 2>/dev/null
ncu -i gpurun_out/prof_bench.ncu-rep --page details > gpurun_out/details_bench.txt 2>/dev/null
rm -f gpurun_out/prof_bench.ncu-rep
