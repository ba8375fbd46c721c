#!/bin/bash
# Round-2 ncu evidence (1 GPU): launch list of the bench command + one `--set full` capture each of the CIPS kernel, the renderer and
# the points_linear kernel.  Numbers printed under ncu are never bench values.
set -u
mkdir -p gpurun_out
O=gpurun_out
python __graft_entry__.py > $O/r02h_build.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/r02h_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-eager > $O/r02h_bench_under_ncu.log 2>&1; echo "launch list: $?"
python tools/launch_list_summary.py $O/r02h_launches.csv > $O/r02h_launch_list_summary.md 2>&1; head -8 $O/r02h_launch_list_summary.md
timeout 600 ncu --set full --clock-control none --import-source on -k regex:cips_tc_kernel -s 1 -c 1 -f -o $O/r02h_cips python tools/prof_cips.py cips 16 > $O/r02h_ncu_cips.log 2>&1; echo "ncu cips: $?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ray_siren_tc_kernel -s 1 -c 1 -f -o $O/r02h_ray python tools/prof_cips.py full 16 > $O/r02h_ncu_ray.log 2>&1; echo "ncu ray: $?"
cat > /tmp/plin_run.py <<'PY'
import sys; sys.path.insert(0, ".")
import torch, cips3d_b200
x = torch.randn(16 * 16384 * 24, 128, device="cuda"); w = torch.randn(128, 128, device="cuda") * 0.1; b = torch.randn(128, device="cuda")
for _ in range(3): y = cips3d_b200.ops._points_linear_raw(x, w, b, None, False)
torch.cuda.synchronize(); print("done", y.shape)
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:plin_kernel -s 1 -c 1 -f -o $O/r02h_plin python /tmp/plin_run.py > $O/r02h_ncu_plin.log 2>&1; echo "ncu plin: $?"
for k in cips ray plin; do
  ncu -i $O/r02h_$k.ncu-rep --page raw --csv > $O/r02h_${k}_raw.csv 2>/dev/null
  python tools/ncu_summary.py $O/r02h_${k}_raw.csv > $O/r02h_ncu_${k}_summary.md 2>&1; head -30 $O/r02h_ncu_${k}_summary.md
  ncu -i $O/r02h_$k.ncu-rep --page source --csv > $O/r02h_${k}_source.csv 2>/dev/null
  python tools/ncu_top.py $O/r02h_${k}_source.csv 25 > $O/r02h_ncu_${k}_top_stalls.txt 2>&1
  rm -f $O/r02h_$k.ncu-rep
done
ls -la $O | grep r02h
