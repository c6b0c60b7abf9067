#!/bin/bash
# One GPU-box round (run under gpurun): smoke, -m gpu parity tests, the default bench, the
# ncu launch list of the bench command and `ncu --set full` captures of the top kernels.
# Everything lands in gpurun_out/.  STAGES selects what runs (default: all).
STAGES=${STAGES:-"smoke tests bench launches full"}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
nproc >> gpurun_out/gpu.txt
for s in $STAGES; do
  case $s in
  smoke)
    echo "== smoke"
    timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -15 | tee gpurun_out/smoke.log
    ;;
  tests)
    echo "== pytest -m gpu"
    timeout ${PYTEST_TIMEOUT:-1500} python -m pytest tests -m gpu -q --tb=short -x --timeout=600 ${PYTEST_ARGS} 2>&1 | tail -${PYTEST_TAIL:-40} | tee gpurun_out/pytest_gpu.log
    ;;
  bench)
    echo "== bench"
    timeout 900 python bench.py ${BENCH_ARGS} 2>&1 | tail -5 | tee gpurun_out/bench.json
    ;;
  refarm)
    echo "== bench --impl reference"
    timeout 900 python bench.py --impl reference --steps 3 --warmup 1 2>&1 | tail -3 | tee gpurun_out/bench_ref.json
    ;;
  launches)
    echo "== ncu launch list"
    timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv \
      --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e ${BENCH_ARGS} \
      > gpurun_out/launches.log 2>&1
    tail -2 gpurun_out/launches.log
    ;;
  full)
    echo "== ncu --set full"
    for k in ${FULL_KERNELS:-k_push_tiled k_fs_scatter k_fs_offsets}; do
      timeout 1200 ncu --set full --clock-control none --import-source on -k regex:$k -s 2 -c 1 -f \
        -o gpurun_out/full_$k python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e ${BENCH_ARGS} \
        > gpurun_out/full_$k.log 2>&1
      tail -2 gpurun_out/full_$k.log
    done
    ;;
  sweep)
    echo "== sweep"
    : > gpurun_out/sweep.jsonl
    while IFS= read -r args; do
      [ -z "$args" ] && continue
      echo "## $args" >> gpurun_out/sweep.jsonl
      timeout 600 python bench.py --no-cpu --no-e2e --steps 5 --warmup 3 $args 2>&1 | tail -2 >> gpurun_out/sweep.jsonl
    done <<< "${SWEEP}"
    ;;
  esac
done
