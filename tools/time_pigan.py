"""pi-GAN renderer timing: fp32-FMA path vs the fused tcgen05 kernel (C3D_PIGAN_IMPL=tc), CUDA events, images/s and the
fraction of the tensor peak on the field's algorithmic FLOPs (526 848 MAC per sample point x 2S points per ray)."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import cips3d_b200
from oracle import cips3d_oracle as O

dev = "cuda:0"
R = int(sys.argv[1]) if len(sys.argv) > 1 else 64
B = int(sys.argv[2]) if len(sys.argv) > 2 else 4
pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
peak = json.load(open(pk))["bf16_tflops"] if os.path.exists(pk) else 1600.0
G = cips3d_b200.ImplicitGenerator3d(cips3d_b200.SPATIALSIRENBASELINE, z_dim=256).to(dev).eval()
G.load_state_dict(O.synthetic_state_dict(O.pigan_template(), seed=3, sigma_bias=0.3))
G.device = G.siren.device = dev
kw = dict(O.PIGAN_KWARGS)
z = torch.randn(B, 256, device=dev)
mac_per_point = 3 * 256 + 7 * 256 * 256 + 256 + 259 * 256 + 3 * 256
flop = 2.0 * mac_per_point * 2 * kw["num_steps"] * R * R * B
# argv[3]: comma-separated subset -- run the hardware-unvalidated tcgen05 forms in their own processes (a protocol error
# traps the CUDA context and would take the remaining timings along)
impls = sys.argv[3].split(",") if len(sys.argv) > 3 else ("simt", "tc", "tc-pair")
for impl in impls:
    os.environ["C3D_PIGAN_IMPL"] = impl.split("-")[0]
    os.environ["C3D_PIGAN_PAIR"] = "1" if impl.endswith("pair") else "0"
    with torch.no_grad():
        for _ in range(3):
            G(z, img_size=R, nerf_noise=0.0, **kw)
        torch.cuda.synchronize()
        ts = []
        for _ in range(5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); G(z, img_size=R, nerf_noise=0.0, **kw); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
    ms = sorted(ts)[len(ts) // 2]
    print(json.dumps(dict(impl=impl, img_size=R, batch=B, ms=ms, img_per_s=B / ms * 1e3, tflops=flop / ms / 1e9, frac_of_tensor_peak=flop / ms / 1e9 / peak)))
