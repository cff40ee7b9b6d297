# A/B timing of alternative builds of the library: bash tools/gpu_ab.sh <suffix> [<suffix> ...]   ("" = product build)
# every build is timed TWICE, interleaved, inside one call (box-to-box variation is ~5 %)
mkdir -p gpurun_out
for round in 1 2; do
for sfx in "$@"; do
  lib=$PWD/dvi_ekf_b200/libeskf_b200$sfx.so
  echo "== $lib"
  ESKF_B200_LIB=$lib timeout 150 python tools/variant_bench.py --variants 3 --n 4096 --reps 3 --stats 2>&1 | tee -a gpurun_out/ab$sfx.log
done
done
