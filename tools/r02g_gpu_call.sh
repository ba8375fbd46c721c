#!/bin/bash
# Round-2 GPU call G: renderer variants on ONE box (box-to-box variance is ~5 %): round-1 kernel vs staged + per-slot issuers
# (+ epilogue turn, + stagger), after the jitter prefetch / loop unrolling.
set -u
mkdir -p gpurun_out
O=gpurun_out
python __graft_entry__.py > $O/r02g_build.log 2>&1
run() { echo "$1: $(env $2 timeout 300 python tools/time_forward.py 16 2>&1 | tail -1 | cut -c1-200)"; }
for rep in 1 2; do
  run "r01 renderer (rep $rep)" "C3D_LIB_PATH=$PWD/cips-3d_b200/libcips3d_b200_r01ray.so"
  run "staged, e_turn 0 (rep $rep)" "C3D_RAY_E_TURN=0"
  run "staged, e_turn 1 (rep $rep)" "C3D_RAY_E_TURN=1"
  run "staged, e_turn 0, stagger 10us (rep $rep)" "C3D_RAY_E_TURN=0 C3D_RAY_STAGGER_NS=10000"
done 2>&1 | tee $O/r02g_ray_variants.txt
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_baseline_sizes_gpu.py -m gpu -q -x -p no:cacheprovider > $O/r02g_pytest_gpu.log 2>&1; echo "gpu tests: exit $?"; tail -2 $O/r02g_pytest_gpu.log
# channels-last discriminator: tests + train step
timeout 400 python -m pytest tests/test_conv2d_cpu.py -q -p no:cacheprovider > $O/r02g_pytest_conv.log 2>&1; echo "conv2d tests: exit $?"
for c in 5 4 3; do
  extra=""; [ $c = 3 ] && extra="--film-backend fused --integrate-backend fused"
  timeout 500 python tools/bench_train_step.py --config $c --cips-backend fused $extra > $O/r02g_train_c$c.json 2> $O/r02g_train_c$c.err; echo "train c$c: $?"; cut -c1-130 $O/r02g_train_c$c.json
done
timeout 500 python tools/bench_train_step.py --config 5 --cips-backend fused --profile $O/r02g_prof_c5.txt > /dev/null 2>&1; head -30 $O/r02g_prof_c5.txt | cut -c1-200
# points_linear (NeRF per-point linears on tcgen05) + the whole default suite
timeout 600 python -m pytest tests/test_film_gpu.py -m gpu -q -x -p no:cacheprovider > $O/r02g_pytest_film.log 2>&1; echo "film/points_linear gpu tests: exit $?"; tail -3 $O/r02g_pytest_film.log
timeout 500 python tools/bench_train_step.py --config 3 --cips-backend fused --film-backend fused --integrate-backend fused --linear-backend fused --profile $O/r02g_prof_c3_linear.txt > $O/r02g_train_c3_linear.json 2> $O/r02g_train_c3_linear.err; echo "train c3 + native linears: $?"; cut -c1-130 $O/r02g_train_c3_linear.json
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider > $O/r02g_pytest_all.log 2>&1; echo "whole gpu suite: exit $?"; tail -3 $O/r02g_pytest_all.log
