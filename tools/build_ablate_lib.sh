#!/bin/bash
# Timing-experiment build of the library: C3D_CIPS_ABLATE=<bits> removes parts of the CIPS kernel's work (see KArgs::ablate).
#   build_ablate_lib.sh [suffix] [extra nvcc flags]    e.g.  build_ablate_lib.sh spin -DC3D_SUSPEND_HINT_NS=-1
set -e
SUF=${1:-}; shift || true
OUT=../libcips3d_b200_ablate${SUF:+_$SUF}.so
cd "$(dirname "$0")/../cips-3d_b200/csrc"
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 --expt-extended-lambda --expt-relaxed-constexpr \
  -Xcompiler -fPIC -shared -DC3D_CIPS_ABLATE "$@" *.cu -o $OUT
echo built $OUT
