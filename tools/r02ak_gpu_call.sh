#!/bin/bash
# Round-2 last GPU call: the full contract line of the final build (both arms) and smoke.
set -u
mkdir -p gpurun_out
O=gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > $O/r02ak_smoke.log 2>&1; echo "smoke: $?"; tail -1 $O/r02ak_smoke.log
timeout 170 python bench.py > $O/r02ak_bench.json 2> $O/r02ak_bench.err; echo "bench: $?"; tail -c 200 $O/r02ak_bench.json; echo
timeout 120 python bench.py --impl reference --steps 3 --warmup 1 > $O/r02ak_bench_reference_arm.json 2> $O/r02ak_bench_reference_arm.err; echo "bench reference arm: $?"; tail -c 200 $O/r02ak_bench_reference_arm.json; echo
