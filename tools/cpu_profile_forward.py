import os, sys, cProfile, pstats, io, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import cips3d_b200
from oracle import cips3d_oracle as O
dev = "cuda:0"
G = cips3d_b200.GeneratorNerfINR(**{k: (dict(v) if isinstance(v, dict) else v) for k, v in O.G_CFG.items()}, device=dev).to(dev).eval()
G.load_state_dict(O.synthetic_state_dict(O.generator_template(), seed=1234))
G.skip_unused_noise_draws = True
zs = G.get_zs(16)
print("torch threads", torch.get_num_threads(), "cpus", os.cpu_count(), flush=True)
with torch.no_grad():
    for _ in range(20):
        G(zs, img_size=256, nerf_noise=0.0, **O.G_KWARGS)
    torch.cuda.synchronize()
    pr = cProfile.Profile()
    pr.enable()
    t0 = time.perf_counter()
    for _ in range(10):
        G(zs, img_size=256, nerf_noise=0.0, **O.G_KWARGS)
    t1 = time.perf_counter()
    pr.disable()
    torch.cuda.synchronize()
print("cpu per step ms", (t1 - t0) * 100)
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(18)
print(s.getvalue()[:3500])
