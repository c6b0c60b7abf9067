#!/bin/bash
# option sweep on one GPU (run under gpurun); results in gpurun_out/sweep.jsonl
mkdir -p gpurun_out
: > gpurun_out/sweep.jsonl
run() { echo "## $*" >> gpurun_out/sweep.jsonl; timeout 600 python bench.py --no-cpu --no-e2e --steps 5 --warmup 3 "$@" 2>&1 | tail -3 >> gpurun_out/sweep.jsonl; }
for args in "$@"; do run $args; done
