#!/bin/bash
# Round-2 GPU call AB: residual loads vs stores vs ToRGB in the pair kernel's epilogue; residual requested a chunk ahead.
set -u
mkdir -p gpurun_out
O=gpurun_out
D=$PWD/cips-3d_b200
for abl in 0 128 512 640 256; do
  echo "pair ablate=$abl: $(C3D_LIB_PATH=$D/libcips3d_b200_ablate.so C3D_CIPS_ABLATE=$abl C3D_CIPS_PAIR=1 timeout 200 python tools/time_cips.py 16 2>&1 | tail -1)"
done 2>&1 | tee $O/r02ab_cips_pair_residual.txt
for rep in 1 2; do
  echo "pair, residual a chunk ahead (rep $rep): $(C3D_LIB_PATH=$D/libcips3d_b200_ablate_early.so C3D_CIPS_PAIR=1 timeout 200 python tools/time_cips.py 16 2>&1 | tail -1)"
done 2>&1 | tee -a $O/r02ab_cips_pair_residual.txt
