#!/bin/bash
# Round-2 GPU call Z: CTA-pair CIPS kernel with N = 256 MMAs and one issuer.
set -u
mkdir -p gpurun_out
O=gpurun_out
D=$PWD/cips-3d_b200
timeout 400 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "cta_pair or umma_pair or cips" -p no:cacheprovider > $O/r02z_pytest.log 2>&1; echo "pair tests: $?"; tail -2 $O/r02z_pytest.log
for rep in 1 2; do
  echo "single (rep $rep): $(timeout 200 python tools/time_cips.py 16 2>&1 | tail -1)"
  echo "pair   (rep $rep): $(C3D_CIPS_PAIR=1 timeout 200 python tools/time_cips.py 16 2>&1 | tail -1)"
done 2>&1 | tee $O/r02z_cips_pair.txt
for abl in 7 23 1 2 4; do
  echo "pair ablate=$abl: $(C3D_LIB_PATH=$D/libcips3d_b200_ablate.so C3D_CIPS_ABLATE=$abl C3D_CIPS_PAIR=1 timeout 200 python tools/time_cips.py 16 2>&1 | tail -1)"
done 2>&1 | tee $O/r02z_cips_pair_ablate.txt
C3D_LIB_PATH=$D/libcips3d_b200_trace_light.so C3D_CIPS_PAIR=1 timeout 200 python tools/trace_cips_light.py 4 4 3 > $O/r02z_cips_light_pair.txt 2>&1; echo "light trace pair: $?"
C3D_CIPS_PAIR=1 timeout 300 python bench.py --no-cpu-baseline --no-eager > $O/r02z_bench_pair.json 2> $O/r02z_bench_pair.err; tail -c 300 $O/r02z_bench_pair.json
