# Micro-benchmarks behind the design decisions, built from their sources on the GPU box:
#   bash tools/gpu_microbench.sh  -> gpurun_out/microbench.log
mkdir -p gpurun_out
out=gpurun_out/microbench.log
: > $out
for name in dmma_peak fp64_lat; do
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/$name tools/microbench/$name.cu >> $out 2>&1 && timeout 120 /tmp/$name >> $out 2>&1 || echo "FAILED $name" >> $out
done
cat $out
