#!/bin/bash
# Round-2 GPU call E: epilogue-turn lock A/B (renderer), separable stream blur.
set -u
mkdir -p gpurun_out
O=gpurun_out
python __graft_entry__.py > $O/r02e_build.log 2>&1
for et in 1 0; do
  C3D_RAY_E_TURN=$et timeout 300 python tools/time_forward.py 16 > $O/r02e_time_forward_eturn$et.log 2>&1; echo "e_turn $et: $(tail -1 $O/r02e_time_forward_eturn$et.log)"
done
C3D_LIB_PATH=$PWD/cips-3d_b200/libcips3d_b200_trace.so timeout 200 python tools/trace_ray.py 16 > $O/r02e_ray_trace_eturn1.txt 2>&1; echo "trace: $?"
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_baseline_sizes_gpu.py -m gpu -q -x -p no:cacheprovider > $O/r02e_pytest_gpu.log 2>&1; echo "gpu tests: exit $?"; tail -3 $O/r02e_pytest_gpu.log
timeout 300 python tools/bench_disc_ops.py 2>&1 | grep blur | cut -c1-200 > $O/r02e_blur.jsonl; cat $O/r02e_blur.jsonl
