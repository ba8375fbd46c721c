#!/bin/bash
# Round-2 GPU call N: is the CIPS weight stream latency-bound (time ~ 1 / ring depth) or bandwidth-bound (independent of depth)?
set -u
mkdir -p gpurun_out
O=gpurun_out
python __graft_entry__.py > $O/r02n_build.log 2>&1
for rep in 1 2; do
  echo "5 stages (rep $rep): $(timeout 200 python tools/time_cips.py 16 2>&1 | tail -1)"
  echo "4 stages (rep $rep): $(C3D_LIB_PATH=$PWD/cips-3d_b200/libcips3d_b200_stages4.so timeout 200 python tools/time_cips.py 16 2>&1 | tail -1)"
  echo "3 stages (rep $rep): $(C3D_LIB_PATH=$PWD/cips-3d_b200/libcips3d_b200_stages3.so timeout 200 python tools/time_cips.py 16 2>&1 | tail -1)"
done 2>&1 | tee $O/r02n_cips_ring_depth.txt
for b in 1 2 4; do echo "5 stages, B=$b: $(timeout 200 python tools/time_cips.py $b 2>&1 | tail -1)"; done 2>&1 | tee -a $O/r02n_cips_ring_depth.txt
