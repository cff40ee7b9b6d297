# ncu --set full capture of the persistent kernel (one launch) + csv exports, read with tools/ncu_digest.py
set -x
mkdir -p gpurun_out
V=${1:-3}; TAG=${2:-v$V}; FRAMES=${3:-40}
export ESKF_B200_VARIANT=$V
timeout 900 ncu --set full --import-source on --clock-control none -k regex:eskf_kernel -s 1 -c 1 -o gpurun_out/prof_$TAG -f python tools/profile_run.py --filters 4096 --frames $FRAMES --passes 2 > gpurun_out/ncu_full_$TAG.log 2>&1; echo rc=$?
tail -3 gpurun_out/ncu_full_$TAG.log
ncu -i gpurun_out/prof_$TAG.ncu-rep --page raw --csv > gpurun_out/raw_$TAG.csv 2>/dev/null
ncu -i gpurun_out/prof_$TAG.ncu-rep --page source --csv > gpurun_out/source_$TAG.csv 2>/dev/null
rm -f gpurun_out/prof_$TAG.ncu-rep
