# ncu launch list (kernel durations) of one variant_bench pass: bash tools/gpu_launchlist.sh <name>
name=$1
mkdir -p gpurun_out
ESKF_B200_LIB=$PWD/dvi_ekf_b200/libeskf_b200_$name.so timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv \
  --log-file gpurun_out/launches_$name.csv python tools/variant_bench.py --variants 3 --n 4096 --reps 1 --stats > gpurun_out/launches_$name.log 2>&1
python - <<PY
import csv
rows=[r for r in csv.reader(open('gpurun_out/launches_$name.csv')) if len(r)>5]
hdr=rows[0]; ik=hdr.index('Kernel Name'); iv=hdr.index('Metric Value')
for r in rows[1:]:
    print(r[ik][:70].ljust(70), r[iv])
PY
