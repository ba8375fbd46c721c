#!/bin/bash
# Round-2 GPU call R: remote mbarrier arrives without the cluster-scope release (MEMBAR.ALL.GPU) -- CTA-pair and multicast-cluster
# CIPS kernels, pi-GAN pair kernel.
set -u
mkdir -p gpurun_out
O=gpurun_out
timeout 400 python -m pytest tests/test_gpu_parity.py tests/test_pigan_gpu.py -m gpu -q -x -k "cta_pair or umma_pair or cips or cluster or pair" -p no:cacheprovider > $O/r02r_pytest.log 2>&1; echo "pair/cluster tests: $?"; tail -2 $O/r02r_pytest.log
for rep in 1 2; do
  echo "single    (rep $rep): $(timeout 200 python tools/time_cips.py 16 2>&1 | tail -1)"
  echo "pair      (rep $rep): $(C3D_CIPS_PAIR=1 timeout 200 python tools/time_cips.py 16 2>&1 | tail -1)"
  echo "cluster 2 (rep $rep): $(C3D_CIPS_CLUSTER=2 timeout 200 python tools/time_cips.py 16 2>&1 | tail -1)"
  echo "cluster 4 (rep $rep): $(C3D_CIPS_CLUSTER=4 timeout 200 python tools/time_cips.py 16 2>&1 | tail -1)"
done 2>&1 | tee $O/r02r_cips_variants.txt
C3D_LIB_PATH=$PWD/cips-3d_b200/libcips3d_b200_trace.so C3D_CIPS_PAIR=1 timeout 200 python tools/trace_cips.py 4 > $O/r02r_cips_trace_pair.txt 2>&1; echo "pair trace: $?"
head -24 $O/r02r_cips_trace_pair.txt | cut -c1-200
grep -A34 "^layer 5, leader" $O/r02r_cips_trace_pair.txt | cut -c1-170
