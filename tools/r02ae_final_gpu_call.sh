#!/bin/bash
# Round-2 final GPU call (1 GPU): smoke, bench (both arms, full contract line), ncu evidence for the final build (launch list of the
# bench command + one --set full capture of the CTA-pair CIPS kernel), one training-step timing (the stash forward shares the
# two-phase epilogue).  Numbers printed under ncu are never bench values.
set -u
mkdir -p gpurun_out
O=gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > $O/r02ae_smoke.log 2>&1; echo "smoke: $?"; tail -2 $O/r02ae_smoke.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:cips_tc_kernel -s 1 -c 1 -f -o $O/r02ae_cips python tools/prof_cips.py cips 16 > $O/r02ae_ncu_cips.log 2>&1; echo "ncu cips: $?"
ncu -i $O/r02ae_cips.ncu-rep --page raw --csv > $O/r02ae_cips_raw.csv 2>/dev/null
python tools/ncu_summary.py $O/r02ae_cips_raw.csv > $O/r02ae_ncu_cips_summary.md 2>&1; head -24 $O/r02ae_ncu_cips_summary.md
ncu -i $O/r02ae_cips.ncu-rep --page source --csv > $O/r02ae_cips_source.csv 2>/dev/null
python tools/ncu_top.py $O/r02ae_cips_source.csv 25 > $O/r02ae_ncu_cips_top_stalls.txt 2>&1
rm -f $O/r02ae_cips.ncu-rep $O/r02ae_cips_source.csv
python tools/update_traffic.py c3d_cips_fwd $O/r02ae_ncu_cips_summary.md profiles/r02ae_ncu_cips_summary.md "B=16, round 2, CTA-pair kernel"; cp profiles/traffic.json $O/r02ae_traffic.json
timeout 500 python bench.py > $O/r02ae_bench.json 2> $O/r02ae_bench.err; echo "bench: $?"; tail -c 300 $O/r02ae_bench.json; echo
timeout 500 python bench.py --impl reference > $O/r02ae_bench_reference_arm.json 2> $O/r02ae_bench_reference_arm.err; echo "bench reference arm: $?"; tail -c 300 $O/r02ae_bench_reference_arm.json; echo
timeout 400 python tools/bench_train_step.py --config 5 --cips-backend fused > $O/r02ae_train_c5.json 2> $O/r02ae_train_c5.err; echo "train c5: $?"; tail -c 400 $O/r02ae_train_c5.json; echo
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/r02ae_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-eager > $O/r02ae_bench_under_ncu.log 2>&1; echo "launch list: $?"
python tools/launch_list_summary.py $O/r02ae_launches.csv > $O/r02ae_launch_list_summary.md 2>&1; head -12 $O/r02ae_launch_list_summary.md
C3D_LIB_PATH=$PWD/cips-3d_b200/libcips3d_b200_trace_light.so timeout 200 python tools/trace_cips_light.py 4 8 3 > $O/r02ae_cips_light_pair_l8.txt 2>&1; echo "light trace: $?"; grep "layer period" $O/r02ae_cips_light_pair_l8.txt | head -1
echo "pigan tc:      $(timeout 200 python tools/time_pigan.py 64 4 2>&1 | tail -1)" | tee $O/r02ae_pigan_pair.txt
echo "pigan tc pair: $(C3D_PIGAN_PAIR=1 timeout 200 python tools/time_pigan.py 64 4 2>&1 | tail -1)" | tee -a $O/r02ae_pigan_pair.txt
ls -la $O | grep r02ae
