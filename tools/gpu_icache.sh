# instruction-supply counters of the persistent kernel, noise on / off (metrics only: a few replays)
M=gpu__time_duration.sum,sm__icc_request_hit_rate.pct,gcc__cache_requests_type_instruction.sum,gcc__cache_requests_type_instruction.sum.pct_of_peak_sustained_elapsed,smsp__inst_executed.sum,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__average_warp_latency_issue_stalled_no_instruction.ratio,smsp__warp_issue_stalled_no_instruction_per_warp_active.pct
mkdir -p gpurun_out
for opt in "" "--no-noise" "--no-noise --no-stats"; do
  tag=$(echo "x$opt" | tr -d ' -')
  timeout 600 ncu --metrics $M --clock-control none -k regex:eskf_kernel -s 1 -c 1 --csv --log-file gpurun_out/icache_$tag.csv \
    python tools/profile_run.py --filters 4096 --frames 0 --passes 2 $opt > gpurun_out/icache_$tag.log 2>&1
  echo "== $opt rc=$?"; grep -v "^==PROF==" gpurun_out/icache_$tag.csv | cut -d, -f13- | tail -12
done
