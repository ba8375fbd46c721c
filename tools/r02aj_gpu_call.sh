#!/bin/bash
# Round-2 GPU call AJ: pair issuer with compile-time steps -- whole GPU suite, kernel timing, light trace, bench.
set -u
mkdir -p gpurun_out
O=gpurun_out
for rep in 1 2 3; do
  echo "pair (rep $rep): $(timeout 200 python tools/time_cips.py 16 2>&1 | tail -1)"
done 2>&1 | tee $O/r02aj_cips_pair.txt
timeout 900 python -m pytest tests -m gpu -q -x -p no:cacheprovider > $O/r02aj_pytest.log 2>&1; echo "gpu suite: $?"; tail -2 $O/r02aj_pytest.log
C3D_LIB_PATH=$PWD/cips-3d_b200/libcips3d_b200_trace_light.so timeout 200 python tools/trace_cips_light.py 4 8 3 > $O/r02aj_cips_light_pair_l8.txt 2>&1; echo "light trace: $?"; grep "layer period" $O/r02aj_cips_light_pair_l8.txt | head -1
timeout 300 python bench.py --no-cpu-baseline --no-eager > $O/r02aj_bench.json 2> $O/r02aj_bench.err; tail -c 200 $O/r02aj_bench.json
