mkdir -p gpurun_out; : > gpurun_out/sweep.jsonl
for lib in libpsc_b200.so libpsc_b200_d2b4.so libpsc_b200_d3b4.so libpsc_b200_d4b5.so libpsc_b200_d4b6.so; do
  echo "## $lib" >> gpurun_out/sweep.jsonl
  PSC_B200_LIB=$lib timeout 600 python bench.py --no-cpu --no-e2e --steps 5 --warmup 3 2>&1 | tail -1 >> gpurun_out/sweep.jsonl
done
