#!/bin/bash
# Round-2 GPU call L: CIPS kernel pipeline traces (single CTA vs cta_group::2 pairs) with the atomic-free stamps; points_linear timing.
set -u
mkdir -p gpurun_out
O=gpurun_out
python __graft_entry__.py > $O/r02l_build.log 2>&1
C3D_LIB_PATH=$PWD/cips-3d_b200/libcips3d_b200_trace.so timeout 200 python tools/trace_cips.py 4 > $O/r02l_cips_trace_default.txt 2>&1; echo "cips trace default: $?"
C3D_LIB_PATH=$PWD/cips-3d_b200/libcips3d_b200_trace.so C3D_CIPS_PAIR=1 timeout 200 python tools/trace_cips.py 4 > $O/r02l_cips_trace_pair.txt 2>&1; echo "cips trace pair: $?"
head -30 $O/r02l_cips_trace_default.txt | cut -c1-200; head -30 $O/r02l_cips_trace_pair.txt | cut -c1-200
python - > $O/r02l_plin_time.txt 2>&1 <<'PY'
import sys; sys.path.insert(0, ".")
import torch, cips3d_b200
for (rows, K, N) in ((16 * 16384 * 24, 128, 128), (16 * 16384 * 24, 128, 64), (16 * 16384 * 24, 64, 32)):
    x = torch.randn(rows, K, device="cuda"); w = torch.randn(N, K, device="cuda") * 0.1; b = torch.randn(N, device="cuda")
    for _ in range(3): y = cips3d_b200.ops._points_linear_raw(x, w, b, None, False)
    torch.cuda.synchronize()
    ts = []
    for _ in range(10):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); y = cips3d_b200.ops._points_linear_raw(x, w, b, None, False); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ms = sorted(ts)[len(ts) // 2]
    gb = rows * (K + N) * 4 / 1e9
    print(f"points_linear rows {rows} K {K} N {N}: {ms:.3f} ms  {gb / ms * 1e3:.0f} GB/s  frac of 6569.6 = {gb / ms * 1e3 / 6569.6:.3f}")
    ref_ms = []
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); yr = torch.nn.functional.linear(x, w, b); e1.record(); torch.cuda.synchronize()
        ref_ms.append(e0.elapsed_time(e1))
    print(f"   torch F.linear (fp32, TF32 off): {sorted(ref_ms)[2]:.3f} ms")
PY
cat $O/r02l_plin_time.txt
