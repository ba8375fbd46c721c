import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
d = int(sys.argv[1]) if len(sys.argv) > 1 else 1
torch.cuda.set_device(d)
import cips3d_b200
from oracle import cips3d_oracle as O
dev = torch.device("cuda", d)
print("device", d, torch.cuda.get_device_name(d), torch.cuda.get_device_properties(d).multi_processor_count, flush=True)
G = cips3d_b200.GeneratorNerfINR(**{k: (dict(v) if isinstance(v, dict) else v) for k, v in O.G_CFG.items()}, device=dev).to(dev).eval()
G.load_state_dict(O.synthetic_state_dict(O.generator_template(), seed=1234))
torch.manual_seed(1000 + d)
zs = {"z_nerf": torch.randn(16, 256).to(dev), "z_inr": torch.randn(16, 512).to(dev)}
with torch.no_grad():
    for i in range(40):
        img, _ = G(zs, img_size=256, nerf_noise=0.0, **O.G_KWARGS)
    torch.cuda.synchronize()
print("device", d, "ok", float(img.abs().mean()), flush=True)
