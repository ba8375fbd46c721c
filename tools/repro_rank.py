import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import cips3d_b200
from oracle import cips3d_oracle as O
dev = "cuda:0"
G = cips3d_b200.GeneratorNerfINR(**{k: (dict(v) if isinstance(v, dict) else v) for k, v in O.G_CFG.items()}, device=dev).to(dev).eval()
G.load_state_dict(O.synthetic_state_dict(O.generator_template(), seed=1234))
for rank in (1, 0, 2, 3):
    torch.manual_seed(1000 + rank)
    zs = {"z_nerf": torch.randn(16, 256).to(dev), "z_inr": torch.randn(16, 512).to(dev)}
    with torch.no_grad():
        for i in range(25):
            img, _ = G(zs, img_size=256, nerf_noise=0.0, **O.G_KWARGS)
            torch.cuda.synchronize()
            if not torch.isfinite(img).all():
                print("rank", rank, "step", i, "non-finite output!", flush=True)
                break
    print("rank seed", rank, "ok", float(img.abs().mean()), flush=True)
