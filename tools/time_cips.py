"""CUDA-event timing of the CIPS kernel alone (B images at r256), median of 10."""
import os, sys
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
    style = G.mapping_network(**zs)
    x = torch.randn(B, 65536, 32, device=dev)
    ws, s1p, dm, rw, rb = G.inr_net.kernel_inputs(style, 9)
    for _ in range(5):
        ops.cips_forward(x, ws, s1p, dm, rw, rb)
    torch.cuda.synchronize()
    ts = []
    for _ in range(10):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); ops.cips_forward(x, ws, s1p, dm, rw, rb); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
ms = sorted(ts)[len(ts) // 2]
print(f"cips B={B}: {ms:.3f} ms  {B * 587.47 / ms:.1f} TFLOP/s  (min {min(ts):.3f})")
