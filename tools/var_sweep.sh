# A/B of library variants built by tools/build_variant.sh (run under gpurun):
#   VARIANTS="libpsc_b200.so libpsc_b200_x.so" BENCH_ARGS="..." bash tools/var_sweep.sh
mkdir -p gpurun_out; : > gpurun_out/sweep.jsonl
for lib in ${VARIANTS:-libpsc_b200.so}; do
  echo "## $lib ${BENCH_ARGS}" >> gpurun_out/sweep.jsonl
  PSC_B200_LIB=$lib timeout 600 python bench.py --no-cpu --no-e2e --steps 5 --warmup 3 ${BENCH_ARGS} 2>&1 | tail -1 >> gpurun_out/sweep.jsonl
done
