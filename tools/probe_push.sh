#!/bin/bash
# compile only the production xyz push kernel (TMA, COUNT, launch bounds 256x3) and print
# its register / spill figures -- quick loop for register-pressure work
cd "$(dirname "$0")/../psc_b200/csrc"
FM=${FM:-false}
nvcc -gencode arch=compute_100a,code=sm_100a -std=c++17 -O3 -lineinfo -Xcompiler -fPIC -I../../include \
  --expt-relaxed-constexpr -Xcudafe --diag_suppress=177 -fmad=$FM -DPUSH_VARIANT=exact -DPUSH_PROBE -DPUSH_PROBE_T=${PT:-256} -DPUSH_PROBE_B=${PB:-3} -DPUSH_PROBE_G=${PG:-true} $PROBE_FLAGS \
  -Xptxas -v -c push.cu -o /tmp/push_probe.o 2> /tmp/probe.log || { cat /tmp/probe.log; exit 1; }
grep -A3 "GeoStatic" /tmp/probe.log | grep "spill\|Used"
