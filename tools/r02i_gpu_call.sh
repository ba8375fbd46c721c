#!/bin/bash
# Round-2 GPU call I: renderer A/B on ONE box -- round-1 kernel vs the current one (staged issue + deferred merge), 3 repetitions.
set -u
mkdir -p gpurun_out
O=gpurun_out
python __graft_entry__.py > $O/r02i_build.log 2>&1
run() { echo "$1: $(env $2 timeout 300 python tools/time_forward.py 16 2>&1 | tail -1 | cut -c1-200)"; }
for rep in 1 2 3; do
  run "r01 renderer (rep $rep)" "C3D_LIB_PATH=$PWD/cips-3d_b200/libcips3d_b200_r01ray.so"
  run "current, e_turn 0 (rep $rep)" "C3D_RAY_E_TURN=0"
  run "current, e_turn 0, stagger 10us (rep $rep)" "C3D_RAY_E_TURN=0 C3D_RAY_STAGGER_NS=10000"
  run "current, e_turn 1 (rep $rep)" "C3D_RAY_E_TURN=1"
done 2>&1 | tee $O/r02i_ray_variants.txt
