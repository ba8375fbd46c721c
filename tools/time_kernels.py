"""Standalone CUDA-event timing of the two fused kernels across batch sizes / cluster sizes."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import cips3d_b200
from cips3d_b200 import ops
from oracle import cips3d_oracle as O

dev = "cuda:0"
G = cips3d_b200.GeneratorNerfINR(**{k: (dict(v) if isinstance(v, dict) else v) for k, v in O.G_CFG.items()}, device=dev).to(dev).eval()
G.load_state_dict(O.synthetic_state_dict(O.generator_template(), seed=1234))
R = 256


def t_cips(B, reps=10):
    zs = G.get_zs(B)
    with torch.no_grad():
        style = G.mapping_network(**zs)
        x = torch.randn(B, R * R, 32, device=dev)
        ws, s1p, dm, rw, rb = G.inr_net.kernel_inputs(style, 9)
        ops.cips_forward(x, ws, s1p, dm, rw, rb)
        torch.cuda.synchronize()
        ts = []
        for _ in range(reps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); ops.cips_forward(x, ws, s1p, dm, rw, rb); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
    print('   reps:', ' '.join(f'{t:.2f}' for t in ts), flush=True)
    return min(ts)


def t_full(B, reps=10):
    zs = G.get_zs(B)
    with torch.no_grad():
        G(zs, img_size=R, nerf_noise=0.0, **O.G_KWARGS)
        torch.cuda.synchronize()
        ops.PROFILE = {}
        for _ in range(reps):
            G(zs, img_size=R, nerf_noise=0.0, **O.G_KWARGS)
        torch.cuda.synchronize()
        prof, ops.PROFILE = ops.PROFILE, None
    for k, v in prof.items():
        print('   ', k, ' '.join(f'{a.elapsed_time(b):.2f}' for a, b in v), flush=True)
    return {k: min(a.elapsed_time(b) for a, b in v) for k, v in prof.items()}


for cl in ("1",):
    os.environ["C3D_CIPS_CLUSTER"] = cl
    for B in (2, 4, 8, 16):
        ms = t_cips(B)
        print(f"cips cluster={cl} B={B}: {ms:.3f} ms  {B*587.47/ms:.1f} TFLOP/s  {B/ms*1e3:.1f} img/s", flush=True)
os.environ["C3D_CIPS_CLUSTER"] = "1"
for B in (2, 8, 16):
    print("full forward B=%d:" % B, t_full(B), flush=True)
