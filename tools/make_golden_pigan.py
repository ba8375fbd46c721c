"""Generate tests/golden/pigan_*.npz from the UNMODIFIED reference's pi-GAN classes (needs /root/reference):
piGAN_lib/generators/generators.py ImplicitGenerator3d + piGAN_lib/siren/siren.py SPATIALSIRENBASELINE / TALLSIREN.

    python tools/make_golden_pigan.py

Stored per case: the weight recipe (numpy PCG64 seed; oracle.synthetic_state_dict over oracle.pigan_template), the
latent, every random draw of the forward in reference order and what the real reference produced (pixels, pitch/yaw,
and the tensors at its fancy_integration call sites).  matplotlib (imported by volumetric_rendering.py, unused on this
path) is stubbed; `curriculums` is the reference's own module."""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)
import ref_capture  # noqa: E402
from oracle import cips3d_oracle as O  # noqa: E402

REF = os.environ.get("CIPS3D_REFERENCE", "/root/reference")
OUT = os.path.join(ROOT, "tests", "golden")


def import_reference_pigan():
    for name in ("matplotlib", "matplotlib.pyplot"):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    lib = os.path.join(REF, "piGAN_lib")
    if lib not in sys.path:
        sys.path.insert(0, lib)
    from generators import generators as G      # piGAN_lib/generators/generators.py
    from generators import volumetric_rendering as VR
    from siren import siren as S                  # piGAN_lib/siren/siren.py
    return G, VR, S


def gen_case(name, *, siren_cls, seed, B, img_size, nerf_noise, z_seed, sigma_bias=0.0, staged_psi=None, **over):
    G, VR, S = import_reference_pigan()
    gen = G.ImplicitGenerator3d(getattr(S, siren_cls), z_dim=256)
    gen.device = gen.siren.device = "cpu"
    tmpl = {k: tuple(v.shape) for k, v in gen.state_dict().items()}
    assert tmpl == {k: tuple(v) for k, v in O.pigan_template().items()}, "state_dict contract changed"
    gen.load_state_dict(O.synthetic_state_dict(tmpl, seed=seed, sigma_bias=sigma_bias))
    kw = dict(O.PIGAN_KWARGS)
    kw.update(over)
    torch.manual_seed(z_seed)
    z = torch.randn(B, 256)
    calls = []
    orig = G.fancy_integration

    def spy(rgb_sigma, z_vals, **k):
        out = orig(rgb_sigma, z_vals, **k)
        calls.append((rgb_sigma.clone(), z_vals.clone(), [o.clone() for o in out]))
        return out

    G.fancy_integration = spy
    log = []
    extra = {}
    try:
        with torch.no_grad(), ref_capture.record_draws(log):
            if staged_psi is None:
                img, py = gen(z, img_size=img_size, nerf_noise=nerf_noise, **kw)
            else:
                # staged_forward draws 10000 latents for the average frequencies first (generators.py:99-107)
                img, depth_map = gen.staged_forward(z, img_size=img_size, nerf_noise=nerf_noise, psi=staged_psi, **kw)
                py = torch.zeros(B, 2)
                extra["avg_frequencies"] = gen.avg_frequencies.numpy()       # mean over the 10000 draws (generators.py:99-107)
                extra["avg_phase_shifts"] = gen.avg_phase_shifts.numpy()
                extra["depth_map"] = depth_map.numpy()
                log = log[1:]
    finally:
        G.fancy_integration = orig
    draws = ref_capture.draws_from_log(log, hierarchical=kw["hierarchical_sample"])
    final = calls[-1]
    out = dict(kwargs_json=np.array(repr(kw)), siren_cls=np.array(siren_cls), seed=seed, sigma_bias=sigma_bias, B=B,
               img_size=img_size, nerf_noise=nerf_noise, staged_psi=-1.0 if staged_psi is None else staged_psi,
               z=z.numpy(), img=img.numpy(), pitch_yaw=py.numpy(), coarse=calls[0][0].numpy(),
               all_z=final[1][..., 0].numpy(), rgb=final[2][0].numpy(), depth=final[2][1][..., 0].numpy(), **extra)
    for k, v in draws.items():
        if v is not None:
            out["draw_" + k] = v.numpy()
    np.savez_compressed(os.path.join(OUT, f"pigan_{name}.npz"), **out)
    print(name, "img", tuple(img.shape), "abs max", float(img.abs().max()))


if __name__ == "__main__":
    gen_case("spatial_r8", siren_cls="SPATIALSIRENBASELINE", seed=71, B=2, img_size=8, nerf_noise=0.0, z_seed=1)
    gen_case("spatial_r8_noise_backs", siren_cls="SPATIALSIRENBASELINE", seed=72, B=2, img_size=8, nerf_noise=0.5, z_seed=2,
             sigma_bias=0.3, clamp_mode="softplus", last_back=True, white_back=True)
    gen_case("tall_r6_lockview", siren_cls="TALLSIREN", seed=73, B=1, img_size=6, nerf_noise=0.0, z_seed=3,
             sigma_bias=0.5, lock_view_dependence=True)
    gen_case("spatial_r6_nohier_s24", siren_cls="SPATIALSIRENBASELINE", seed=74, B=2, img_size=6, nerf_noise=0.0, z_seed=4,
             sigma_bias=0.5, hierarchical_sample=False, num_steps=24)
    gen_case("spatial_r6_staged_psi07", siren_cls="SPATIALSIRENBASELINE", seed=75, B=1, img_size=6, nerf_noise=0.0, z_seed=5,
             sigma_bias=0.4, staged_psi=0.7)
