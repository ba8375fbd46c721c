#!/bin/bash
# Round-2 GPU call B: ray-kernel phase traces (default / warp math), train-step kernel profiles, the new reference arm.
set -u
mkdir -p gpurun_out
O=gpurun_out
python __graft_entry__.py > $O/r02b_build.log 2>&1
C3D_LIB_PATH=$PWD/cips-3d_b200/libcips3d_b200_trace.so timeout 200 python tools/trace_ray.py 16 > $O/r02b_ray_trace_default.txt 2>&1; echo "trace default: $?"
C3D_LIB_PATH=$PWD/cips-3d_b200/libcips3d_b200_trace.so C3D_RAY_MATH=warp timeout 200 python tools/trace_ray.py 16 > $O/r02b_ray_trace_warp.txt 2>&1; echo "trace warp: $?"
C3D_LIB_PATH=$PWD/cips-3d_b200/libcips3d_b200_trace.so C3D_RAY_MATH=fold timeout 200 python tools/trace_ray.py 16 > $O/r02b_ray_trace_fold.txt 2>&1; echo "trace fold: $?"
C3D_LIB_PATH=$PWD/cips-3d_b200/libcips3d_b200_trace.so timeout 200 python tools/trace_cips.py > $O/r02b_cips_trace_default.txt 2>&1; echo "cips trace default: $?"
C3D_LIB_PATH=$PWD/cips-3d_b200/libcips3d_b200_trace.so C3D_CIPS_PAIR=1 timeout 200 python tools/trace_cips.py > $O/r02b_cips_trace_pair.txt 2>&1; echo "cips trace pair: $?"
timeout 500 python tools/bench_train_step.py --config 5 --cips-backend fused --profile $O/r02b_prof_c5_fused.txt > $O/r02b_train_c5_fused.json 2> $O/r02b_train_c5_fused.err; echo "train c5 fused: $?"
timeout 500 python tools/bench_train_step.py --config 3 --film-backend fused --integrate-backend fused --profile $O/r02b_prof_c3.txt > $O/r02b_train_c3.json 2> $O/r02b_train_c3.err; echo "train c3: $?"
timeout 500 python tools/bench_train_step.py --config 4 --cips-backend fused --profile $O/r02b_prof_c4.txt > $O/r02b_train_c4.json 2> $O/r02b_train_c4.err; echo "train c4: $?"
timeout 400 python bench.py --impl reference --steps 5 --warmup 1 > $O/r02b_bench_reference.json 2> $O/r02b_bench_reference.err; echo "bench reference arm: $?"
C3D_STYLE_PREP=fused timeout 600 python bench.py --steps 20 --warmup 5 > $O/r02b_bench.json 2> $O/r02b_bench.err; echo "bench: $?"
cat $O/r02b_train_c*.json | cut -c1-300
cat $O/r02b_bench_reference.json $O/r02b_bench.json | cut -c1-2500
head -60 $O/r02b_prof_c5_fused.txt | cut -c1-200
