#!/bin/bash
# Round-2 GPU call AC: two-phase epilogue (hand the chunk over before the residual stores / ToRGB) vs the previous build.
set -u
mkdir -p gpurun_out
O=gpurun_out
D=$PWD/cips-3d_b200
timeout 400 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "cta_pair or umma_pair or cips" -p no:cacheprovider > $O/r02ac_pytest.log 2>&1; echo "cips tests: $?"; tail -2 $O/r02ac_pytest.log
for rep in 1 2; do
  echo "pair,   one-phase epilogue (rep $rep): $(C3D_LIB_PATH=$D/libcips3d_b200_epi1.so C3D_CIPS_PAIR=1 timeout 200 python tools/time_cips.py 16 2>&1 | tail -1)"
  echo "pair,   two-phase epilogue (rep $rep): $(C3D_CIPS_PAIR=1 timeout 200 python tools/time_cips.py 16 2>&1 | tail -1)"
  echo "single, one-phase epilogue (rep $rep): $(C3D_LIB_PATH=$D/libcips3d_b200_epi1.so timeout 200 python tools/time_cips.py 16 2>&1 | tail -1)"
  echo "single, two-phase epilogue (rep $rep): $(timeout 200 python tools/time_cips.py 16 2>&1 | tail -1)"
done 2>&1 | tee $O/r02ac_cips_epilogue.txt
C3D_LIB_PATH=$D/libcips3d_b200_trace_light.so C3D_CIPS_PAIR=1 timeout 200 python tools/trace_cips_light.py 4 8 3 > $O/r02ac_cips_light_pair_l8.txt 2>&1; echo "light trace pair: $?"
C3D_CIPS_PAIR=1 timeout 300 python bench.py --no-cpu-baseline --no-eager > $O/r02ac_bench_pair.json 2> $O/r02ac_bench_pair.err; tail -c 250 $O/r02ac_bench_pair.json
