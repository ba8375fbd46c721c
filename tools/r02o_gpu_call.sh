#!/bin/bash
set -u
mkdir -p gpurun_out
O=gpurun_out
python __graft_entry__.py > $O/r02o_build.log 2>&1
C3D_LIB_PATH=$PWD/cips-3d_b200/libcips3d_b200_trace.so C3D_CIPS_PAIR=1 timeout 200 python tools/trace_cips.py 4 > $O/r02o_cips_trace_pair.txt 2>&1; echo "pair trace: $?"
C3D_LIB_PATH=$PWD/cips-3d_b200/libcips3d_b200_trace.so timeout 200 python tools/trace_cips.py 4 > $O/r02o_cips_trace_default.txt 2>&1; echo "default trace: $?"
grep -A80 "^layer 5, leader" $O/r02o_cips_trace_pair.txt | cut -c1-170
