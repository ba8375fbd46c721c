#!/bin/bash
# debug build of the library with in-kernel clock stamps (-DC3D_TRACE): cips-3d_b200/libcips3d_b200_trace.so (git-ignored, travels)
set -e
cd "$(dirname "$0")/../cips-3d_b200/csrc"
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 --expt-extended-lambda --expt-relaxed-constexpr \
  -Xcompiler -fPIC -shared -DC3D_TRACE *.cu -o ../libcips3d_b200_trace.so
echo built ../libcips3d_b200_trace.so
