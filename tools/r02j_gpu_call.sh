#!/bin/bash
# Round-2 GPU call J: renderer A/B on one box -- round-1 kernel vs form 0 (polling issuer + round-2 prefetch / deferred merge) vs form 1.
set -u
mkdir -p gpurun_out
O=gpurun_out
python __graft_entry__.py > $O/r02j_build.log 2>&1
run() { echo "$1: $(env $2 timeout 300 python tools/time_forward.py 16 2>&1 | tail -1 | cut -c1-200)"; }
for rep in 1 2 3; do
  run "r01 renderer (rep $rep)" "C3D_LIB_PATH=$PWD/cips-3d_b200/libcips3d_b200_r01ray.so"
  run "form 0 (rep $rep)" "C3D_RAY_SCHED=0"
  run "form 0, stagger 10us (rep $rep)" "C3D_RAY_SCHED=0 C3D_RAY_STAGGER_NS=10000"
  run "form 1, stagger 10us (rep $rep)" "C3D_RAY_SCHED=1 C3D_RAY_STAGGER_NS=10000"
done 2>&1 | tee $O/r02j_ray_variants.txt
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_baseline_sizes_gpu.py -m gpu -q -x -p no:cacheprovider > $O/r02j_pytest_gpu.log 2>&1; echo "gpu tests: exit $?"; tail -2 $O/r02j_pytest_gpu.log
C3D_LIB_PATH=$PWD/cips-3d_b200/libcips3d_b200_trace.so timeout 200 python tools/trace_ray.py 16 > $O/r02j_ray_trace_form0.txt 2>&1; echo "trace: $?"
