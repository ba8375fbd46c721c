"""Phase timeline of two ray-group iterations of the fused renderer, stamped by lane 0 of every worker warp and by the MMA
issuer (needs the -DC3D_TRACE build: `bash tools/build_trace_lib.sh`, then C3D_LIB_PATH=cips-3d_b200/libcips3d_b200_trace.so).
Prints, per slot and phase, the earliest and latest warp (clocks relative to the first event) and the issuer's MMA issues."""
import os, sys, ctypes as C, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import cips3d_b200
from cips3d_b200 import _lib
from oracle import cips3d_oracle as O
dev = "cuda:0"
B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
G = cips3d_b200.GeneratorNerfINR(**{k: (dict(v) if isinstance(v, dict) else v) for k, v in O.G_CFG.items()}, device=dev).to(dev).eval()
G.load_state_dict(O.synthetic_state_dict(O.generator_template(), seed=1234))
zs = G.get_zs(B)
lib = _lib.load()
lib.c3d_debug_ray_trace.argtypes = [C.c_void_p, C.c_int]
buf = (C.c_ulonglong * 4096)()
with torch.no_grad():
    for _ in range(3):
        G(zs, img_size=256, nerf_noise=0.0, **O.G_KWARGS)
    torch.cuda.synchronize()
    lib.c3d_debug_ray_trace(buf, 4096)
    G(zs, img_size=256, nerf_noise=0.0, **O.G_KWARGS)
    n = lib.c3d_debug_ray_trace(buf, 4096)
ev = []
for v in buf[:n]:
    t, tag, a0 = v & 0xFFFFFFFFFF, (v >> 56) & 0xFF, (v >> 40) & 0xFFFF
    ev.append((t, tag, a0 >> 15, (a0 >> 12) & 7, (a0 >> 8) & 1, a0 & 0xFF))
ev.sort()
t0 = ev[0][0]
names = {14: "MMA done (spin)", 1: "MMA issue", 2: "A0 written", 3: "D0 ready", 4: "E0 done", 5: "D1 ready", 6: "E1 done", 7: "D2 ready", 8: "E2 done",
         9: "D3 ready", 10: "E3+sync", 11: "resample done", 12: "merge done", 13: "composite done"}
print("events", n, "math mode", lib.c3d_debug_ray_math_mode())
groups = collections.OrderedDict()
for t, tag, sl, tw, itp, ph in ev:
    if tag == 1:
        groups[(t, "issuer", sl, itp, ph)] = [t]
    else:
        groups.setdefault(("w", sl, itp, ph, tag), []).append(t)
rows = []
for k, ts in groups.items():
    if k[1] == "issuer":
        rows.append((ts[0], f"slot {k[2]} it{k[3]}  MMA issue layer-phase #{k[4]}"))
    else:
        _, sl, itp, ph, tag = k
        rows.append((min(ts), f"slot {sl} it{itp}  {names.get(tag, tag):16s} (#{ph:2d})  first warp {min(ts) - t0:7d}  last warp {max(ts) - t0:7d}  skew {max(ts) - min(ts):5d}  warps {len(ts)}"))
rows.sort()
for t, s in rows:
    print(f"{t - t0:8d}  {s}")
