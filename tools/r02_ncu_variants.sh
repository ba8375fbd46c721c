#!/bin/bash
# Round 2, second GPU call (only for the variants that passed r02_first_gpu_call.sh): one `ncu --set full` capture per
# kernel variant at the bench workload, brought back as .ncu-rep + raw CSV for tools/ncu_summary.py.
#   /usr/local/graft/bin/gpurun --timeout 1500 -- 'bash tools/r02_ncu_variants.sh'
# ncu replays each kernel ~40 times: ONE launch per capture (-c 1), one GPU, never a bench number from these runs.
set -u
mkdir -p gpurun_out
O=gpurun_out
cap() {   # cap <tag> <kernel regex> <env...> -- <command...>
  local tag=$1 rx=$2; shift 2
  local envs=()
  while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
  env "${envs[@]}" timeout 600 ncu --set full --clock-control none --import-source on -k "regex:$rx" -c 1 -o $O/r02b_$tag -f "$@" > $O/r02b_$tag.log 2>&1
  ncu -i $O/r02b_$tag.ncu-rep --page raw --csv > $O/r02b_${tag}_raw.csv 2>/dev/null
  python tools/ncu_summary.py $O/r02b_${tag}_raw.csv "" > $O/r02b_${tag}_summary.md 2>/dev/null
  echo "$tag: $(grep -m1 'gpu__time_duration' $O/r02b_${tag}_summary.md)"
}
cap cips_default  cips_tc_kernel      C3D_X=0            -- python tools/prof_cips.py cips 16 256
cap cips_pair     cips_tc_kernel      C3D_CIPS_PAIR=1    -- python tools/prof_cips.py cips 16 256
cap ray_default   ray_siren_tc_kernel C3D_X=0            -- python tools/prof_cips.py ray 16 256
cap ray_warpmath  ray_siren_tc_kernel C3D_RAY_MATH=warp  -- python tools/prof_cips.py ray 16 256
cap ray_foldmath  ray_siren_tc_kernel C3D_RAY_MATH=fold  -- python tools/prof_cips.py ray 16 256
# HBM-bound kernels: dram bytes vs the algorithmic bytes in DESIGN.md (5th launch of each = past the warm-up)
capskip() {   # like cap, but skips the first launches of the kernel
  local tag=$1 rx=$2; shift 2
  local envs=()
  while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
  env "${envs[@]}" timeout 600 ncu --set full --clock-control none --import-source on -k "regex:$rx" --launch-skip 4 -c 1 -o $O/r02b_$tag -f "$@" > $O/r02b_$tag.log 2>&1
  ncu -i $O/r02b_$tag.ncu-rep --page raw --csv > $O/r02b_${tag}_raw.csv 2>/dev/null
  python tools/ncu_summary.py $O/r02b_${tag}_raw.csv "" > $O/r02b_${tag}_summary.md 2>/dev/null
  echo "$tag: $(grep -m1 'gpu__time_duration' $O/r02b_${tag}_summary.md)"
}
capskip adam_ema   adam_ema_kernel        C3D_X=0         -- python tools/bench_optim.py
capskip blur_tma   blur_tma_kernel        C3D_BLUR_TMA=1  -- python tools/bench_disc_ops.py
capskip image_u8   image_u8_flat4_kernel  C3D_X=0         -- python tools/bench_disc_ops.py
capskip film_bwd   film_sin_bwd_kernel    C3D_X=0         -- python tools/bench_disc_ops.py
capskip integ_fwd  integrate_fwd_kernel   C3D_X=0         -- python tools/bench_disc_ops.py
capskip integ_bwd  integrate_bwd_kernel   C3D_X=0         -- python tools/bench_disc_ops.py
