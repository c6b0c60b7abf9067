#!/bin/bash
# Compile-checks include/psc_b200/psc_adapters_b200.hxx against the reference's own headers
# where they lie (REF, default /root/reference/src).  Prints "skipped" when the reference
# tree is not present (the GPU box).  Nothing is linked or run.
set -e
HERE=$(cd "$(dirname "$0")" && pwd)
ROOT=$HERE/../..
REF=${REF:-/root/reference/src}
if [ ! -d "$REF/include" ]; then
  echo "check_adapters: skipped (no reference tree at $REF)"
  exit 0
fi
${CXX:-g++} -std=c++17 -fsyntax-only -I "$HERE/psc_shim" -I "$ROOT/oracle/shim" -I "$ROOT/include" \
  -I "$REF/include" -I "$REF/kg/include" -I "$REF/libmrc/include" "$HERE/check_adapters.cxx"
echo "check_adapters: MparticlesB200Psc compiles against $REF/include/particles.hxx, particles_simple.hxx"
