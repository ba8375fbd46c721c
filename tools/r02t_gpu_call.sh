#!/bin/bash
# Round-2 GPU call T: where the CIPS kernel's ~11.6 k clk/layer of pure synchronisation (r02s: ablate=7) goes -- wait flavour
# (suspend hint / default try_wait / spin), commit vs plain arrive, light trace of the layer boundary.
set -u
mkdir -p gpurun_out
O=gpurun_out
D=$PWD/cips-3d_b200
for lib in ablate ablate_nohint ablate_spin; do
  for abl in 0 7 15; do
    echo "$lib single ablate=$abl: $(C3D_LIB_PATH=$D/libcips3d_b200_$lib.so C3D_CIPS_ABLATE=$abl timeout 200 python tools/time_cips.py 16 2>&1 | tail -1)"
  done
  for abl in 0 7; do
    echo "$lib pair   ablate=$abl: $(C3D_LIB_PATH=$D/libcips3d_b200_$lib.so C3D_CIPS_ABLATE=$abl C3D_CIPS_PAIR=1 timeout 200 python tools/time_cips.py 16 2>&1 | tail -1)"
  done
done 2>&1 | tee $O/r02t_cips_waits.txt
for abl in 0 7 15; do
  C3D_LIB_PATH=$D/libcips3d_b200_trace_light.so C3D_CIPS_ABLATE=$abl timeout 200 python tools/trace_cips_light.py 4 4 3 > $O/r02t_cips_light_single_abl$abl.txt 2>&1; echo "light trace abl=$abl: $?"
done
C3D_LIB_PATH=$D/libcips3d_b200_trace_light.so C3D_CIPS_ABLATE=0 C3D_CIPS_PAIR=1 timeout 200 python tools/trace_cips_light.py 4 4 3 > $O/r02t_cips_light_pair_abl0.txt 2>&1; echo "light trace pair: $?"
