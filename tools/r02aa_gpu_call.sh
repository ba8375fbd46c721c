#!/bin/bash
# Round-2 GPU call AA: which part of the epilogue slows the pair kernel's MMAs (TMEM loads, A-operand stores, residual traffic, ToRGB).
set -u
mkdir -p gpurun_out
O=gpurun_out
D=$PWD/cips-3d_b200
for abl in 0 32 64 96 128 256 384 480; do
  echo "pair ablate=$abl: $(C3D_LIB_PATH=$D/libcips3d_b200_ablate.so C3D_CIPS_ABLATE=$abl C3D_CIPS_PAIR=1 timeout 200 python tools/time_cips.py 16 2>&1 | tail -1)"
done 2>&1 | tee $O/r02aa_cips_pair_epilogue_ablate.txt
C3D_LIB_PATH=$D/libcips3d_b200_trace_light.so C3D_CIPS_PAIR=1 timeout 200 python tools/trace_cips_light.py 4 8 3 > $O/r02aa_cips_light_pair_l8.txt 2>&1; echo "light trace pair: $?"
