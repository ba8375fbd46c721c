"""Run the CIPS / ray kernels once on synthetic inputs (for ncu captures)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import cips3d_b200
from oracle import cips3d_oracle as O

which = sys.argv[1] if len(sys.argv) > 1 else "cips"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 4
R = int(sys.argv[3]) if len(sys.argv) > 3 else 256
dev = "cuda:0"
G = cips3d_b200.GeneratorNerfINR(**{k: (dict(v) if isinstance(v, dict) else v) for k, v in O.G_CFG.items()}, device=dev).to(dev).eval()
G.load_state_dict(O.synthetic_state_dict(O.generator_template(), seed=1234))
zs = G.get_zs(B)
with torch.no_grad():
    for _ in range(2):
        if which == "cips":
            style = G.mapping_network(**zs)
            x = torch.randn(B, R * R, 32, device=dev)
            out = G.inr_net(x, style)
        else:
            out, _ = G(zs, img_size=R, nerf_noise=0.0, **O.G_KWARGS)
torch.cuda.synchronize()
print("done", tuple(out.shape))
