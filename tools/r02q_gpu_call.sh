#!/bin/bash
# Round-2 GPU call Q: CTA-pair CIPS kernel with the peer relay merged into the leader's full barrier.
set -u
mkdir -p gpurun_out
O=gpurun_out
python __graft_entry__.py > $O/r02q_build.log 2>&1
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "cta_pair or umma_pair or cips" -p no:cacheprovider > $O/r02q_pytest.log 2>&1; echo "pair tests: $?"; tail -2 $O/r02q_pytest.log
for rep in 1 2; do
  echo "single (rep $rep): $(timeout 200 python tools/time_cips.py 16 2>&1 | tail -1)"
  echo "pair   (rep $rep): $(C3D_CIPS_PAIR=1 timeout 200 python tools/time_cips.py 16 2>&1 | tail -1)"
done 2>&1 | tee $O/r02q_cips_pair.txt
C3D_LIB_PATH=$PWD/cips-3d_b200/libcips3d_b200_trace.so C3D_CIPS_PAIR=1 timeout 200 python tools/trace_cips.py 4 > $O/r02q_cips_trace_pair.txt 2>&1; echo "pair trace: $?"
head -24 $O/r02q_cips_trace_pair.txt | cut -c1-200
grep -A34 "^layer 5, leader" $O/r02q_cips_trace_pair.txt | cut -c1-170
