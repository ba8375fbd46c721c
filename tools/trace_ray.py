"""Phase timeline of one ray-group iteration of the fused renderer (needs the -DC3D_TRACE build)."""
import os, sys, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import cips3d_b200
from cips3d_b200 import _lib
from oracle import cips3d_oracle as O
dev = "cuda:0"
G = cips3d_b200.GeneratorNerfINR(**{k: (dict(v) if isinstance(v, dict) else v) for k, v in O.G_CFG.items()}, device=dev).to(dev).eval()
G.load_state_dict(O.synthetic_state_dict(O.generator_template(), seed=1234))
zs = G.get_zs(4)
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
ev = sorted([((v & 0xFFFFFFFFFF), (v >> 56) & 0xFF, ((v >> 40) & 0xFFFF) >> 8, (v >> 40) & 0xFF) for v in buf[:n]])
t0 = ev[0][0]
names = {1: "MMA issue", 2: "A0 written", 3: "D0 ready", 4: "E0 done", 5: "D1 ready", 6: "E1 done", 7: "D2 ready", 8: "E2 done",
         9: "D3 ready", 10: "E3+sync", 11: "resample done", 12: "merge done", 13: "composite done"}
print("events", n)
for t, tag, sl, ph in ev:
    print(f"{t - t0:8d}  slot {sl}  {names.get(tag, tag)}  (#{ph})")
