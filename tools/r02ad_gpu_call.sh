#!/bin/bash
# Round-2 GPU call AD: the whole -m gpu suite with the final defaults (CTA-pair CIPS kernel, fp16 residual stream for image-only calls),
# CIPS kernel timing with the fp16 / fp32 residual stream, bench.
set -u
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -p no:cacheprovider > $O/r02ad_pytest.log 2>&1; echo "gpu suite: $?"; tail -3 $O/r02ad_pytest.log
for rep in 1 2; do
  echo "pair, fp16 residual (rep $rep): $(timeout 200 python tools/time_cips.py 16 2>&1 | tail -1)"
  echo "pair, fp32 residual (rep $rep): $(C3D_CIPS_RES16=0 timeout 200 python tools/time_cips.py 16 2>&1 | tail -1)"
done 2>&1 | tee $O/r02ad_cips_res16.txt
echo "single, fp16 residual: $(C3D_CIPS_PAIR=0 timeout 200 python tools/time_cips.py 16 2>&1 | tail -1)" | tee -a $O/r02ad_cips_res16.txt
timeout 300 python bench.py --no-cpu-baseline --no-eager > $O/r02ad_bench.json 2> $O/r02ad_bench.err; tail -c 250 $O/r02ad_bench.json
