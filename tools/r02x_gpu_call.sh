#!/bin/bash
# Round-2 GPU call X: cost of an MMA-issuer step (microbenchmark); ring depth of the two-ring pair kernel (3 vs 5 stages per ring).
set -u
mkdir -p gpurun_out
O=gpurun_out
D=$PWD/cips-3d_b200
timeout 120 tools/ubench/issue_bench 2>&1 | tee $O/r02x_issue_bench.txt
for lib in ablate ablate_st3; do
  for abl in 0 7 23; do
    echo "$lib pair ablate=$abl: $(C3D_LIB_PATH=$D/libcips3d_b200_$lib.so C3D_CIPS_ABLATE=$abl C3D_CIPS_PAIR=1 timeout 200 python tools/time_cips.py 16 2>&1 | tail -1)"
  done
done 2>&1 | tee $O/r02x_cips_pair_depth.txt
