#!/bin/bash
# Round-2 GPU call C: per-slot issuers (timing + trace), cluster-occupancy diagnostic, stream blur, ops.conv2d in the train step.
set -u
mkdir -p gpurun_out
O=gpurun_out
python __graft_entry__.py > $O/r02c_build.log 2>&1
python - > $O/r02c_cluster_diag.txt 2>&1 <<'PY'
import sys; sys.path.insert(0, ".")
import cips3d_b200
lib = cips3d_b200._lib.load()
for cl, pair in ((1, 0), (2, 0), (2, 1)):
    print("max active clusters cl=%d pair=%d:" % (cl, pair), lib.c3d_debug_cips_max_clusters(cl, pair), lib.c3d_last_error())
PY
cat $O/r02c_cluster_diag.txt
timeout 900 python -m pytest tests -m gpu -q -x -p no:cacheprovider > $O/r02c_pytest_gpu.log 2>&1; echo "gpu suite: exit $?"; tail -3 $O/r02c_pytest_gpu.log
timeout 300 python tools/time_forward.py 16 > $O/r02c_time_forward.log 2>&1; tail -1 $O/r02c_time_forward.log
C3D_RAY_MATH=warp timeout 300 python tools/time_forward.py 16 > $O/r02c_time_forward_warp.log 2>&1; tail -1 $O/r02c_time_forward_warp.log
C3D_LIB_PATH=$PWD/cips-3d_b200/libcips3d_b200_trace.so timeout 200 python tools/trace_ray.py 16 > $O/r02c_ray_trace_default.txt 2>&1; echo "trace: $?"
timeout 300 python tools/bench_disc_ops.py > $O/r02c_disc_ops.jsonl 2>&1; grep blur $O/r02c_disc_ops.jsonl | cut -c1-200
C3D_BLUR=tile timeout 300 python tools/bench_disc_ops.py 2>&1 | grep blur | cut -c1-200 > $O/r02c_disc_ops_tile.jsonl
timeout 500 python tools/bench_train_step.py --config 5 --cips-backend fused --profile $O/r02c_prof_c5.txt > $O/r02c_train_c5.json 2> $O/r02c_train_c5.err; echo "train c5: $?"
timeout 500 python tools/bench_train_step.py --config 3 --cips-backend fused --film-backend fused --integrate-backend fused --profile $O/r02c_prof_c3.txt > $O/r02c_train_c3.json 2> $O/r02c_train_c3.err; echo "train c3: $?"
timeout 500 python tools/bench_train_step.py --config 4 --cips-backend fused > $O/r02c_train_c4.json 2> $O/r02c_train_c4.err; echo "train c4: $?"
cat $O/r02c_train_c*.json | cut -c1-200
timeout 600 python bench.py --steps 20 --warmup 5 > $O/r02c_bench.json 2> $O/r02c_bench.err; echo "bench: $?"
cat $O/r02c_bench.json | cut -c1-4000
head -40 $O/r02c_prof_c5.txt | cut -c1-200
