"""Generate tests/golden/*.npz from the UNMODIFIED reference (needs /root/reference).

    python tools/make_golden.py

Each fixture stores: the weight recipe (numpy PCG64 seed + sigma_bias; see
oracle.cips3d_oracle.synthetic_state_dict), the latents, every random draw of the forward
in reference order, and what the *real* reference produced: final images, pitch/yaw and
the tensors seen at its two fancy_integration call sites (coarse sigma/features, merged z,
integrated 32-d pixel features).  Nothing here runs on the GPU box; the fixtures do.
"""
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

OUT = os.path.join(ROOT, "tests", "golden")


def gen_case(G, name, *, seed, sigma_bias, B, img_size, nerf_noise, z_seed, **over):
    from exp.pigan import pigan_utils
    tmpl = {k: tuple(v.shape) for k, v in G.state_dict().items()}
    G.load_state_dict(O.synthetic_state_dict(tmpl, seed=seed, sigma_bias=sigma_bias))
    kw = dict(ref_shim.G_KWARGS)
    kw.update(over)
    torch.manual_seed(z_seed)
    zs = G.get_zs(B)
    calls = []
    orig = pigan_utils.fancy_integration

    def spy(rgb_sigma, z_vals, **k):
        out = orig(rgb_sigma=rgb_sigma, z_vals=z_vals, **k)
        calls.append((rgb_sigma.clone(), z_vals.clone(), [o.clone() for o in out]))
        return out

    pigan_utils.fancy_integration = spy
    log = []
    try:
        with torch.no_grad(), ref_capture.record_draws(log):
            img, py = G(zs, img_size=img_size, nerf_noise=nerf_noise, return_aux_img=True, **kw)
    finally:
        pigan_utils.fancy_integration = orig
    hier = kw["hierarchical_sample"]
    d = ref_capture.draws_from_log(log, hierarchical=hier)
    rec = dict(
        seed=seed, sigma_bias=sigma_bias, B=B, img_size=img_size, nerf_noise=nerf_noise,
        kwargs_json=np.array(repr(kw)),
        z_nerf=zs["z_nerf"].numpy(), z_inr=zs["z_inr"].numpy(),
        img=img.numpy(), pitch_yaw=py.numpy(),
        pixels_fea=calls[-1][2][0].numpy(), depth=calls[-1][2][1][..., 0].numpy(),
        all_z=calls[-1][1][..., 0].numpy(),
        coarse=calls[0][0].numpy() if hier else calls[-1][0].numpy(),
    )
    for k, v in d.items():
        rec["draw_" + k] = v.numpy()
    path = os.path.join(OUT, f"gen_{name}.npz")
    np.savez_compressed(path, **rec)
    print(name, "img", tuple(img.shape), "absmean %.4f" % img.abs().mean().item(),
          "%.0f KB" % (os.path.getsize(path) / 1024))


def gen_disc(D, name, *, seed, R, B, alpha, aux):
    tmpl = {k: tuple(v.shape) for k, v in D.state_dict().items()}
    D.load_state_dict(O.synthetic_state_dict(tmpl, seed=seed))
    rng = np.random.Generator(np.random.PCG64(seed + 1))
    x = torch.from_numpy(rng.uniform(-1, 1, (B, 3, R, R)).astype(np.float32))
    with torch.no_grad():
        out = D(x, use_aux_disc=aux, alpha=alpha)[0]
    path = os.path.join(OUT, f"disc_{name}.npz")
    np.savez_compressed(path, seed=seed, R=R, B=B, alpha=alpha, aux=aux, x=x.numpy(), out=out.numpy())
    print(name, out.flatten().tolist())


def main():
    os.makedirs(OUT, exist_ok=True)
    torch.manual_seed(1234)
    G = ref_shim.build_reference_generator().eval()
    if "r8_hier_s24" in sys.argv[1:]:
        # the web-demo sampling (ffhq_exp.yaml:244-278: num_steps 24, hierarchical G_kwargs -> 24 coarse + 24 fine = 48 sorted
        # samples per ray), added in round 2 without regenerating the other fixtures
        gen_case(G, "r8_hier_s24", seed=79, sigma_bias=0.25, B=2, img_size=8, nerf_noise=0.0, z_seed=11, num_steps=24)
        return
    tmpl = {k: list(v.shape) for k, v in G.state_dict().items()}
    # state_dict contract (key order + shapes) of the reference G and D
    D = ref_shim.build_reference_discriminator().eval()
    import json
    with open(os.path.join(OUT, "state_dict_contract.json"), "w") as f:
        json.dump(dict(generator=tmpl, discriminator={k: list(v.shape) for k, v in D.state_dict().items()}), f)
    gen_case(G, "r16_synth", seed=1234, sigma_bias=0.0, B=2, img_size=16, nerf_noise=0.0, z_seed=7)
    gen_case(G, "r16_trained_noise", seed=4321, sigma_bias=0.4, B=2, img_size=16, nerf_noise=0.5, z_seed=8)
    gen_case(G, "r8_softplus_backs", seed=77, sigma_bias=0.2, B=3, img_size=8, nerf_noise=0.25, z_seed=9,
             clamp_mode="softplus", last_back=True, white_back=True)
    gen_case(G, "r8_nohier_s24", seed=78, sigma_bias=0.2, B=2, img_size=8, nerf_noise=0.0, z_seed=10,
             hierarchical_sample=False, num_steps=24, ray_start=0.8, ray_end=1.2, h_stddev=0.5, v_stddev=0.4)
    gen_disc(D, "r32_main", seed=99, R=32, B=4, alpha=1.0, aux=False)
    gen_disc(D, "r64_aux_fade", seed=100, R=64, B=4, alpha=0.3, aux=True)
    # ray-index contract: pixel (h, w) <-> ray h*W + w  (comm_utils.py:392-395)
    from exp.comm import comm_utils
    pts, z, d = comm_utils.get_initial_rays_trig(bs=1, num_steps=3, fov=12, resolution=(5, 5),
                                                 ray_start=0.88, ray_end=1.12, device="cpu")
    np.savez_compressed(os.path.join(OUT, "rays_r5.npz"), dirs=d[0].numpy(), z=z[0, 0, :, 0].numpy())


if __name__ == "__main__":
    main()
