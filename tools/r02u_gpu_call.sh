#!/bin/bash
# Round-2 GPU call U: CIPS kernel, who probes the mbarriers (every lane / lane 0 / one lane + named barrier) and the weight ring alone.
set -u
mkdir -p gpurun_out
O=gpurun_out
D=$PWD/cips-3d_b200
for lib in ablate_poll0 ablate_poll1 ablate_poll2; do
  for abl in 0 7 23 18 16; do
    echo "$lib single ablate=$abl: $(C3D_LIB_PATH=$D/libcips3d_b200_$lib.so C3D_CIPS_ABLATE=$abl timeout 200 python tools/time_cips.py 16 2>&1 | tail -1)"
  done
  for abl in 0 7; do
    echo "$lib pair   ablate=$abl: $(C3D_LIB_PATH=$D/libcips3d_b200_$lib.so C3D_CIPS_ABLATE=$abl C3D_CIPS_PAIR=1 timeout 200 python tools/time_cips.py 16 2>&1 | tail -1)"
  done
done 2>&1 | tee $O/r02u_cips_poll.txt
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "cips" -p no:cacheprovider > $O/r02u_pytest.log 2>&1; echo "cips tests (lane-0 probes): $?"; tail -2 $O/r02u_pytest.log
