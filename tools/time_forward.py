"""Where does a full generator forward spend its time? (CUDA events: whole step vs the two fused kernels)"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import cips3d_b200
from cips3d_b200 import ops
from oracle import cips3d_oracle as O
dev = "cuda:0"
B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
G = cips3d_b200.GeneratorNerfINR(**{k: (dict(v) if isinstance(v, dict) else v) for k, v in O.G_CFG.items()}, device=dev).to(dev).eval()
G.load_state_dict(O.synthetic_state_dict(O.generator_template(), seed=1234))
zs = G.get_zs(B)
with torch.no_grad():
    for _ in range(40):
        G(zs, img_size=256, nerf_noise=0.0, **O.G_KWARGS)      # long warm-up: SM clocks ramp from idle slowly
torch.cuda.synchronize()
for skip in (True, True, True):
    G.skip_unused_noise_draws = skip
    with torch.no_grad():
        for _ in range(3):
            G(zs, img_size=256, nerf_noise=0.0, **O.G_KWARGS)
        torch.cuda.synchronize()
        ops.PROFILE = {}
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter(); e0.record()
        n = 10
        for _ in range(n):
            G(zs, img_size=256, nerf_noise=0.0, **O.G_KWARGS)
        e1.record(); t_cpu = time.perf_counter() - t0
        torch.cuda.synchronize()
        prof, ops.PROFILE = ops.PROFILE, None
    ms = e0.elapsed_time(e1) / n
    k = {kk: sum(a.elapsed_time(b) for a, b in v) / n for kk, v in prof.items()}
    for kk, v in prof.items():
        print("    ", kk, " ".join(f"{a.elapsed_time(b):.1f}" for a, b in v), flush=True)
    print(f"skip_noise_draws={skip}: step {ms:.2f} ms  ({B/ms*1e3:.0f} img/s)  cpu-enqueue {t_cpu/n*1e3:.2f} ms  kernels {k}  other {ms - sum(k.values()):.2f} ms", flush=True)
