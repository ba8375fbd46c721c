"""Generate tests/golden/points_forward.npz from the UNMODIFIED reference (needs /root/reference; tools/ref_shim.py):
comm_utils.get_world_points_and_direction (comm_utils.py:682-763) -> GeneratorNerfINR.points_forward (generator.py:1659-1762)
on the ffhq G config with oracle.synthetic_state_dict weights (numpy PCG64: platform-independent bytes).

    python tools/make_golden_points.py

Stored per case: latents, EVERY torch.rand / torch.randn draw in call order (so a replacement can replay them on any device),
the seven outputs of get_world_points_and_direction, the subset of rays, and the real method's inr / aux images."""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)
import ref_capture  # noqa: E402
import ref_shim  # noqa: E402
from oracle import cips3d_oracle as O  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden", "points_forward.npz")
# name: (img_size, hierarchical, nerf_noise, clamp, white_back, last_back, every-other-ray subset)
CASES = {"hier_noise_lastback": (8, True, 0.4, "relu", False, True, True),
         "flat_softplus_white": (6, False, 0.0, "softplus", True, False, False)}


def main():
    ref_shim.install()
    from exp.comm import comm_utils as ref_cu
    G = ref_shim.build_reference_generator().eval()
    G.load_state_dict(O.synthetic_state_dict(O.generator_template(), seed=77, sigma_bias=0.3))
    out = {}
    for idx, (name, (R, hier, noise, clamp, wb, lb, subset)) in enumerate(CASES.items()):
        rng = np.random.Generator(np.random.PCG64(9100 + idx))
        zs = {"z_nerf": torch.from_numpy(rng.standard_normal((2, 256)).astype(np.float32)),
              "z_inr": torch.from_numpy(rng.standard_normal((2, 512)).astype(np.float32))}
        log = []
        torch.manual_seed(300 + idx)
        with torch.no_grad(), ref_capture.record_draws(log):
            world = ref_cu.get_world_points_and_direction(
                batch_size=2, num_steps=12, img_size=R, fov=12, ray_start=0.88, ray_end=1.12, h_stddev=0.3, v_stddev=0.155,
                h_mean=np.pi * 0.5, v_mean=np.pi * 0.5, sample_dist="gaussian", lock_view_dependence=False, device="cpu")
            pts, dirs_exp, origins, dirs, z_vals, pitch, yaw = world
            n_world = len(log)
            idx_grad = torch.arange(0, R * R, 2) if subset else None
            style = G.mapping_network(**zs)
            inr, aux = G.points_forward(
                style_dict=style, transformed_points=pts.view(2, R * R, 12, 3),
                transformed_ray_directions_expanded=dirs_exp.view(2, R * R, 12, 3), num_steps=12, hierarchical_sample=hier,
                z_vals=z_vals, clamp_mode=clamp, nerf_noise=noise, transformed_ray_origins=origins,
                transformed_ray_directions=dirs, white_back=wb, last_back=lb, return_aux_img=True, idx_grad=idx_grad)
        for k, v in zs.items():
            out[f"{name}/{k}"] = v.numpy()
        for i, (kind, t) in enumerate(log):
            out[f"{name}/draw{i:02d}_{kind}"] = t.numpy()
        for k, v in zip(("points", "dirs_exp", "origins", "dirs", "z_vals", "pitch", "yaw"), world):
            out[f"{name}/world_{k}"] = v.numpy()
        out[f"{name}/inr"], out[f"{name}/aux"] = inr.numpy(), aux.numpy()
        out[f"{name}/cfg"] = np.array([R, int(hier), int(clamp == "softplus"), int(wb), int(lb), int(subset), n_world], dtype=np.int64)
        out[f"{name}/nerf_noise"] = np.float32(noise)
    np.savez_compressed(OUT, **out)
    print("wrote", OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
