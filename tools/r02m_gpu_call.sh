#!/bin/bash
# Round-2 GPU call M: CIPS kernel with a per-CTA start stagger (decorrelate the 148 identical weight streams), same box.
set -u
mkdir -p gpurun_out
O=gpurun_out
python __graft_entry__.py > $O/r02m_build.log 2>&1
for rep in 1 2; do for st in 0 20 50 100 200 500 2000; do
  echo "stagger ${st} ns/CTA (rep $rep): $(C3D_CIPS_STAGGER_NS=$st timeout 200 python tools/time_cips.py 16 2>&1 | tail -1)"
done; done 2>&1 | tee $O/r02m_cips_stagger.txt
for st in 0 200; do echo "pair, stagger ${st}: $(C3D_CIPS_PAIR=1 C3D_CIPS_STAGGER_NS=$st timeout 200 python tools/time_cips.py 16 2>&1 | tail -1)"; done 2>&1 | tee -a $O/r02m_cips_stagger.txt
