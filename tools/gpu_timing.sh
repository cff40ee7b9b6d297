# phase timing of profiling builds: bash tools/gpu_timing.sh <tag> <name> ...   -> gpurun_out/timing_<tag>.log
tag=$1; shift
mkdir -p gpurun_out
out=gpurun_out/timing_$tag.log
: > $out
for name in "$@"; do
  lib=$PWD/dvi_ekf_b200/libeskf_b200_$name.so
  ESKF_B200_LIB=$lib timeout 60 python tools/timing_run.py >> $out 2>&1 || echo "FAILED rc=$? ($name)" >> $out
  ESKF_B200_LIB=$lib timeout 60 python tools/timing_run.py --noise --stats >> $out 2>&1 || echo "FAILED rc=$? ($name)" >> $out
done
cat $out
