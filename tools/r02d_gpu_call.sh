#!/bin/bash
# Round-2 GPU call D: ray kernel with staged MMA issue + per-slot issuers: timing, stagger A/B, low-overhead trace; integrate tests.
set -u
mkdir -p gpurun_out
O=gpurun_out
python __graft_entry__.py > $O/r02d_build.log 2>&1
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_baseline_sizes_gpu.py tests/test_integrate_gpu.py -m gpu -q -x -p no:cacheprovider > $O/r02d_pytest_gpu.log 2>&1; echo "gpu tests: exit $?"; tail -3 $O/r02d_pytest_gpu.log
for st in 0 3000 6000 10000; do
  C3D_RAY_STAGGER_NS=$st timeout 300 python tools/time_forward.py 16 > $O/r02d_time_forward_stagger$st.log 2>&1; echo "stagger $st: $(tail -1 $O/r02d_time_forward_stagger$st.log)"
done
C3D_LIB_PATH=$PWD/cips-3d_b200/libcips3d_b200_trace.so timeout 200 python tools/trace_ray.py 16 > $O/r02d_ray_trace.txt 2>&1; echo "trace: $?"
C3D_LIB_PATH=$PWD/cips-3d_b200/libcips3d_b200_trace.so C3D_RAY_STAGGER_NS=6000 timeout 200 python tools/trace_ray.py 16 > $O/r02d_ray_trace_stagger6000.txt 2>&1; echo "trace stagger: $?"
