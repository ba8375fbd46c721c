#!/bin/bash
# Debug build of the library with in-kernel clock stamps (tools/trace_ray.py, tools/trace_cips.py, tools/trace_cips_light.py).
#   build_trace_lib.sh [level] [extra nvcc flags]     level 2 (default) = every stamp; 1 = "light" (CIPS: a few tiles per layer,
#   one epilogue warp -- leaves the traced CTA's timing close to the untraced one).  Output: libcips3d_b200_trace[_light].so
set -e
LEVEL=${1:-2}; shift || true
OUT=../libcips3d_b200_trace.so
[ "$LEVEL" = 1 ] && OUT=../libcips3d_b200_trace_light.so
cd "$(dirname "$0")/../cips-3d_b200/csrc"
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 --expt-extended-lambda --expt-relaxed-constexpr \
  -Xcompiler -fPIC -shared -DC3D_TRACE=$LEVEL "$@" *.cu -o $OUT
echo built $OUT
