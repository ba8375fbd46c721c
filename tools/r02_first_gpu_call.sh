#!/bin/bash
# Round-2 opening GPU call (one gpurun, ~25 GPU-minutes; every item has its own timeout):
#   /usr/local/graft/bin/gpurun --timeout 2400 -- 'bash tools/r02_first_gpu_call.sh'
# 1. the default GPU suite (must stay green), 2. the hardware-unvalidated variants that passed the CPU emulation, each
# in its own process under `timeout` (a protocol error traps the context: it must not take the other checks along),
# 3. A/B timings of every variant against the default, 4. roofline benches of the HBM-bound ops.  Everything lands in
# gpurun_out/r02a_*.  Never a number from under a profiler.
set -u
mkdir -p gpurun_out
O=gpurun_out
python __graft_entry__.py > $O/r02a_build.log 2>&1
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > $O/r02a_pytest_gpu.log 2>&1; echo "validated gpu suite: exit $?" | tee $O/r02a_summary.txt
# first hardware run of the rows built on the emulator (tests/conftest.py HW_FIRST_RUN_FILES): hard failures here, one process per file
for f in film inference integrate optim; do
  C3D_HW_STRICT=1 timeout 300 python -m pytest tests/test_${f}_gpu.py -m gpu -q -rA > $O/r02a_pytest_first_$f.log 2>&1; echo "first hardware run $f: exit $?" | tee -a $O/r02a_summary.txt
done
for k in "umma_pair_selftest" "cips_cta_pair" "blur_tma" "warp_per_ray" "fold_math" "cips_backward_chain"; do
  C3D_EXPERIMENTAL=1 timeout 300 python -m pytest tests/test_gpu_parity.py -q -k "$k" > $O/r02a_pytest_$k.log 2>&1
  echo "experimental $k: exit $?" | tee -a $O/r02a_summary.txt
done
# A/B: CIPS kernel alone (B = 4, 16) and the whole forward
timeout 300 python tools/time_kernels.py > $O/r02a_time_kernels_default.log 2>&1
C3D_CIPS_PAIR=1 timeout 300 python tools/time_kernels.py > $O/r02a_time_kernels_pair.log 2>&1; echo "pair timing: exit $?" | tee -a $O/r02a_summary.txt
timeout 300 python tools/time_forward.py 16 > $O/r02a_time_forward_default.log 2>&1
C3D_CIPS_PAIR=1 timeout 300 python tools/time_forward.py 16 > $O/r02a_time_forward_pair.log 2>&1
C3D_STYLE_PREP=fused timeout 300 python tools/time_forward.py 16 > $O/r02a_time_forward_styleprep.log 2>&1
C3D_RAY_MATH=warp timeout 300 python tools/time_forward.py 16 > $O/r02a_time_forward_raywarp.log 2>&1; echo "ray warp-math timing: exit $?" | tee -a $O/r02a_summary.txt
C3D_RAY_MATH=fold timeout 300 python tools/time_forward.py 16 > $O/r02a_time_forward_rayfold.log 2>&1; echo "ray fold-math timing: exit $?" | tee -a $O/r02a_summary.txt
# HBM-bound ops
timeout 300 python tools/bench_disc_ops.py > $O/r02a_disc_ops_default.jsonl 2>&1
C3D_BLUR_TMA=1 timeout 300 python tools/bench_disc_ops.py > $O/r02a_disc_ops_blur_tma.jsonl 2>&1; echo "blur_tma bench: exit $?" | tee -a $O/r02a_summary.txt
C3D_HW_STRICT=1 C3D_EXPERIMENTAL=1 timeout 300 python -m pytest tests/test_pigan_gpu.py -q -rA > $O/r02a_pytest_pigan.log 2>&1; echo "pigan (simt + tc): exit $?" | tee -a $O/r02a_summary.txt
for impl in simt tc tc-pair; do
  timeout 300 python tools/time_pigan.py 64 4 $impl >> $O/r02a_time_pigan.jsonl 2>> $O/r02a_time_pigan.err; echo "pigan timing $impl: exit $?" | tee -a $O/r02a_summary.txt
done
timeout 300 python tools/bench_optim.py > $O/r02a_optim.jsonl 2>&1; echo "optim bench: exit $?" | tee -a $O/r02a_summary.txt
# train-step configurations of BASELINE.json (3: r128 non-frozen + aux, 5: r256 finetune recipe), fused vs torch optimiser tail
for c in 3 5; do for o in fused torch; do
  timeout 600 python tools/bench_train_step.py --config $c --optim $o > $O/r02a_train_c${c}_$o.json 2> $O/r02a_train_c${c}_$o.err; echo "train step config $c optim $o: exit $?" | tee -a $O/r02a_summary.txt
done; done
timeout 600 python tools/bench_train_step.py --config 3 --optim fused --film-backend fused > $O/r02a_train_c3_fused_film.json 2> $O/r02a_train_c3_fused_film.err
timeout 600 python tools/bench_train_step.py --config 3 --optim fused --film-backend fused --integrate-backend fused > $O/r02a_train_c3_fused_film_integ.json 2> $O/r02a_train_c3_fused_film_integ.err
timeout 600 python tools/bench_train_step.py --config 5 --optim fused --cips-backend fused > $O/r02a_train_c5_fused_cipsbwd.json 2> $O/r02a_train_c5_fused_cipsbwd.err
timeout 600 python tools/bench_train_step.py --config 5 --optim fused --tf32 > $O/r02a_train_c5_fused_tf32.json 2> $O/r02a_train_c5_fused_tf32.err
python bench.py --steps 20 --warmup 5 --u8 > $O/r02a_bench.json 2> $O/r02a_bench.err
# the same contract line with the variants that passed above (only meaningful if their tests exited 0)
C3D_CIPS_PAIR=1 C3D_RAY_MATH=warp timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-eager > $O/r02a_bench_variants.json 2> $O/r02a_bench_variants.err
tail -n 3 $O/r02a_time_kernels_default.log $O/r02a_time_kernels_pair.log $O/r02a_time_forward_default.log $O/r02a_time_forward_pair.log $O/r02a_time_forward_raywarp.log $O/r02a_time_forward_rayfold.log $O/r02a_time_forward_styleprep.log
grep -h "blur" $O/r02a_disc_ops_default.jsonl $O/r02a_disc_ops_blur_tma.jsonl | cut -c1-200
cat $O/r02a_optim.jsonl $O/r02a_time_pigan.jsonl | cut -c1-300
cat $O/r02a_train_c*.json 2>/dev/null | cut -c1-400
cat $O/r02a_summary.txt
python tools/r02_report.py $O > $O/r02a_report.md 2>&1   # copy to profiles/r02a_summary.md after reading it
