# A/B timing of library variants built by tools/build_variants.py:  bash tools/gpu_ab2.sh <tag> <name> [<name> ...]
# Every build is timed in two interleaved rounds inside ONE gpurun call (box-to-box variation is ~5 %); the digest of the final
# state shows which builds are bit-identical.  Output: gpurun_out/ab2_<tag>.log
tag=$1; shift
mkdir -p gpurun_out
out=gpurun_out/ab2_$tag.log
: > $out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm --format=csv,noheader >> $out
for round in 1 2; do
for name in "$@"; do
  lib=$PWD/dvi_ekf_b200/libeskf_b200_$name.so
  echo "== round $round $name" >> $out
  ESKF_B200_LIB=$lib timeout 60 python tools/variant_bench.py --variants 3 --n 4096 --reps 3 --stats --digest >> $out 2>&1 || echo "FAILED rc=$? ($name)" >> $out
done
done
cat $out
