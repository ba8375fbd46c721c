#!/bin/bash
# Round-2 GPU call S: ablation timing of the CIPS kernel (which of epilogue / MMA / weight loads paces a layer), single and pair;
# bench with the pair kernel.
set -u
mkdir -p gpurun_out
O=gpurun_out
A=$PWD/cips-3d_b200/libcips3d_b200_ablate.so
for abl in 0 1 2 4 3 5 6 7; do
  echo "single ablate=$abl: $(C3D_LIB_PATH=$A C3D_CIPS_ABLATE=$abl timeout 200 python tools/time_cips.py 16 2>&1 | tail -1)"
  echo "pair   ablate=$abl: $(C3D_LIB_PATH=$A C3D_CIPS_ABLATE=$abl C3D_CIPS_PAIR=1 timeout 200 python tools/time_cips.py 16 2>&1 | tail -1)"
done 2>&1 | tee $O/r02s_cips_ablate.txt
timeout 300 python bench.py --no-cpu-baseline --no-eager > $O/r02s_bench_single.json 2> $O/r02s_bench_single.err; tail -c 600 $O/r02s_bench_single.json
C3D_CIPS_PAIR=1 timeout 300 python bench.py --no-cpu-baseline --no-eager > $O/r02s_bench_pair.json 2> $O/r02s_bench_pair.err; tail -c 600 $O/r02s_bench_pair.json
