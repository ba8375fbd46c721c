#!/bin/bash
# Round-2 GPU call K: sanity of the restored renderer against the round-1 library on one box, whole GPU suite, bench (both arms),
# then the ncu evidence (tools/r02h_ncu_call.sh).
set -u
mkdir -p gpurun_out
O=gpurun_out
python __graft_entry__.py > $O/r02k_build.log 2>&1
run() { echo "$1: $(env $2 timeout 300 python tools/time_forward.py 16 2>&1 | tail -1 | cut -c1-200)"; }
for rep in 1 2; do
  run "r01 library (rep $rep)" "C3D_LIB_PATH=$PWD/cips-3d_b200/libcips3d_b200_r01ray.so"
  run "this build (rep $rep)" "C3D_X=1"
done 2>&1 | tee $O/r02k_ray_sanity.txt
timeout 1200 python -m pytest tests -m gpu -q -rA -p no:cacheprovider > $O/r02k_pytest_gpu.log 2>&1; echo "whole gpu suite: exit $?"; tail -3 $O/r02k_pytest_gpu.log; grep -h "TRAIN-PARITY" $O/r02k_pytest_gpu.log | sort -u | cut -c1-300
timeout 400 python bench.py --impl reference --steps 5 --warmup 1 > $O/r02k_bench_reference.json 2> $O/r02k_bench_reference.err; echo "bench reference arm: $?"
timeout 900 python bench.py --steps 20 --warmup 5 > $O/r02k_bench.json 2> $O/r02k_bench.err; echo "bench: $?"; cut -c1-3000 $O/r02k_bench.json; tail -3 $O/r02k_bench.err
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/r02k_smoke.log 2>&1; echo "smoke: $?"; tail -2 $O/r02k_smoke.log
bash tools/r02h_ncu_call.sh
