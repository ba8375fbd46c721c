#!/bin/bash
# Round-2 GPU call A: parity holes first, then A/B timings of every emulation-verified variant.
#   /usr/local/graft/bin/gpurun --timeout 1800 -- 'bash tools/r02a_gpu_call.sh'
set -u
mkdir -p gpurun_out
O=gpurun_out
S=$O/r02a_summary.txt
: > $S
rm -f $O/r02_parity_baseline_sizes.jsonl
python __graft_entry__.py > $O/r02a_build.log 2>&1
# 1. whole default suite, everything a hard failure
timeout 900 python -m pytest tests -m gpu -q -rA -p no:cacheprovider > $O/r02a_pytest_gpu.log 2>&1; echo "default gpu suite: exit $?" | tee -a $S
grep -h "PARITY\|RANGE" $O/r02a_pytest_gpu.log | sort -u > $O/r02a_parity_lines.txt
# 2. hardware-unvalidated variants, one process each
for k in "umma_pair_selftest" "cips_cta_pair" "blur_tma" "warp_per_ray" "fold_math" "cips_backward_chain"; do
  C3D_EXPERIMENTAL=1 timeout 300 python -m pytest tests/test_gpu_parity.py -q -rA -k "$k" > $O/r02a_pytest_$k.log 2>&1
  echo "experimental $k: exit $?" | tee -a $S
done
C3D_EXPERIMENTAL=1 timeout 300 python -m pytest tests/test_pigan_gpu.py -q -rA > $O/r02a_pytest_pigan.log 2>&1; echo "pigan (simt + tc): exit $?" | tee -a $S
# 3. A/B timings
timeout 300 python tools/time_kernels.py > $O/r02a_time_kernels_default.log 2>&1
C3D_CIPS_PAIR=1 timeout 300 python tools/time_kernels.py > $O/r02a_time_kernels_pair.log 2>&1; echo "pair timing: exit $?" | tee -a $S
timeout 300 python tools/time_forward.py 16 > $O/r02a_time_forward_default.log 2>&1
C3D_STYLE_PREP=fused timeout 300 python tools/time_forward.py 16 > $O/r02a_time_forward_styleprep.log 2>&1
C3D_RAY_MATH=warp timeout 300 python tools/time_forward.py 16 > $O/r02a_time_forward_raywarp.log 2>&1; echo "ray warp-math timing: exit $?" | tee -a $S
C3D_RAY_MATH=fold timeout 300 python tools/time_forward.py 16 > $O/r02a_time_forward_rayfold.log 2>&1; echo "ray fold-math timing: exit $?" | tee -a $S
timeout 300 python tools/bench_disc_ops.py > $O/r02a_disc_ops_default.jsonl 2>&1
C3D_BLUR_TMA=1 timeout 300 python tools/bench_disc_ops.py > $O/r02a_disc_ops_blur_tma.jsonl 2>&1; echo "blur_tma bench: exit $?" | tee -a $S
for impl in simt tc tc-pair; do
  timeout 200 python tools/time_pigan.py 64 4 $impl >> $O/r02a_time_pigan.jsonl 2>> $O/r02a_time_pigan.err; echo "pigan timing $impl: exit $?" | tee -a $S
done
timeout 200 python tools/bench_optim.py > $O/r02a_optim.jsonl 2>&1; echo "optim bench: exit $?" | tee -a $S
# 4. first train-step numbers (config 5: the freeze-NeRF finetune recipe; torch graph vs native CIPS backward)
timeout 400 python tools/bench_train_step.py --config 5 --optim fused > $O/r02a_train_c5_fused.json 2> $O/r02a_train_c5_fused.err; echo "train c5 fused-optim: exit $?" | tee -a $S
timeout 400 python tools/bench_train_step.py --config 5 --optim fused --cips-backend fused > $O/r02a_train_c5_fused_cipsbwd.json 2> $O/r02a_train_c5_fused_cipsbwd.err; echo "train c5 cips-bwd: exit $?" | tee -a $S
timeout 400 python tools/bench_train_step.py --config 3 --optim fused > $O/r02a_train_c3_fused.json 2> $O/r02a_train_c3_fused.err; echo "train c3: exit $?" | tee -a $S
tail -n 3 $O/r02a_time_kernels_default.log $O/r02a_time_kernels_pair.log $O/r02a_time_forward_default.log $O/r02a_time_forward_raywarp.log $O/r02a_time_forward_rayfold.log $O/r02a_time_forward_styleprep.log
grep -h "blur" $O/r02a_disc_ops_default.jsonl $O/r02a_disc_ops_blur_tma.jsonl | cut -c1-200
cat $O/r02a_optim.jsonl $O/r02a_time_pigan.jsonl | cut -c1-300
cat $O/r02a_train_c*.json 2>/dev/null | cut -c1-400
cat $O/r02a_parity_lines.txt | cut -c1-600
tail -n 5 $O/r02a_pytest_gpu.log
cat $S
