#!/bin/bash
# CPU stand-ins for compute-sanitizer on the emulated kernels:
#   memcheck  -> AddressSanitizer build of the emulated library: out-of-bounds global / shared / workspace accesses of any
#                kernel (torch's CPU allocations and the emulator's shared memory carry ASan redzones under LD_PRELOAD)
#   alignment -> -fsanitize=alignment in trap mode: a misaligned float4 / uint4 / uint64 access (a fault on the GPU, silently
#                tolerated by x86) executes ud2 and kills the run
# Usage: bash tools/emu_sanitize.sh [pytest args]      (defaults to the emulation suites; ~15 min)
set -u
cd "$(dirname "$0")/.."
ARGS=("$@")
[ ${#ARGS[@]} -eq 0 ] && ARGS=(tests/test_emu_cpu.py tests/test_optim_cpu.py tests/test_image_export_cpu.py)
ASAN=$(/usr/bin/gcc -print-file-name=libasan.so)
echo "== alignment"
# the fault-injection tests are excluded: the emulator's error path unwinds across fiber stacks, which the sanitizer runtime
# does not survive
SKIP="not detects and not reports_broken"
C3D_EMU_SANITIZE=alignment python -m pytest "${ARGS[@]}" -x -q -k "$SKIP" 2>&1 | tail -3
echo "== address"
LD_PRELOAD=$(readlink -f "$ASAN") ASAN_OPTIONS=detect_leaks=0:detect_stack_use_after_return=0:halt_on_error=1 \
  C3D_EMU_SANITIZE=address python -m pytest "${ARGS[@]}" -x -q -k "$SKIP" 2>&1 | tail -30
