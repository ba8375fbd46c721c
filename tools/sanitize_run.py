"""Tiny generator forward for compute-sanitizer runs."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import cips3d_b200
from oracle import cips3d_oracle as O
dev = "cuda:0"
G = cips3d_b200.GeneratorNerfINR(**{k: (dict(v) if isinstance(v, dict) else v) for k, v in O.G_CFG.items()}, device=dev).to(dev).eval()
G.load_state_dict(O.synthetic_state_dict(O.generator_template(), seed=1234))
R = int(sys.argv[1]) if len(sys.argv) > 1 else 32
zs = G.get_zs(3)
with torch.no_grad():
    img, _ = G(zs, img_size=R, nerf_noise=0.3, **O.G_KWARGS)
    img2, _ = G(zs, img_size=R, nerf_noise=0.0, return_aux_img=True, **{**O.G_KWARGS, "hierarchical_sample": False})
torch.cuda.synchronize()
print("ok", tuple(img.shape), float(img.abs().mean()), tuple(img2.shape))
