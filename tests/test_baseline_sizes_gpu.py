"""GPU parity at the BASELINE.json geometries: the CUDA path through the public class surface
(GeneratorNerfINR.forward -> C-ABI) against the CPU oracle on identical weights, latents and replayed
random draws, at FFHQ r64 B=4 (config 2 geometry), r128 B=2 (config 3) and r256 B=1 (configs 4/5 and the
bench geometry).  Tolerance = north_star's 1e-3 relative fp32 on the rendered RGB, stated per pixel:

  * every pixel whose error exceeds 1e-3 of the image scale must be a ray on which the REFERENCE ITSELF is a
    step function: delta_last = 1e10 (pigan_utils.py:243) turns the last sorted sample's alpha into
    1 - exp(-1e10 * relu(sigma_last)), i.e. 0 for sigma_last <= 0 and 1 for sigma_last > ~1e-9.  The mask
    is |sigma_last| < SIGMA_EPS * max|sigma| with sigma_last taken from the ORACLE (SIGMA_EPS = 1e-3: five
    times the measured pre-integration max-rel error bound of 2e-4, tests/test_gpu_parity.py);
  * the observed outlier fraction, max-rel and l2-rel are printed (pytest -rA) and bounded.

The oracle finishes r256 B=1 in a few seconds on the GPU box's host cores.
"""
import json
import os

import pytest
import torch

from oracle import cips3d_oracle as O
from _util import build_generator, draws_sequence, rel_err, replay_draws

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
SIGMA_EPS = 1e-3
TOL = 1e-3


@pytest.fixture(scope="module")
def pkg():
    import __graft_entry__ as g
    g.build()
    import cips3d_b200
    return cips3d_b200


def baseline_case(R, B, sigma_bias, seed=77, dev=None):
    """-> dict of stats comparing G.forward and ops.render_features with the oracle (no asserts)."""
    dev = dev or DEV
    sd = O.synthetic_state_dict(O.generator_template(), seed=seed, sigma_bias=sigma_bias)
    G = build_generator(dev, sd)
    kw = dict(O.G_KWARGS)
    S = kw["num_steps"]
    g = torch.Generator().manual_seed(1000 * R + B)
    zs = {"z_nerf": torch.randn(B, 256, generator=g), "z_inr": torch.randn(B, 512, generator=g)}
    draws = O.draw_randoms(B, R, S, generator=g)
    with torch.no_grad():
        ref_img, ref_py, r = O.generator_forward(sd, zs, draws, img_size=R, nerf_noise=0.0, return_all=True, **kw)
        zs_d = {k: v.to(dev) for k, v in zs.items()}
        with replay_draws(draws_sequence(draws, True), dev):
            img, py = G(zs_d, img_size=R, nerf_noise=0.0, **kw)
        # the renderer alone, for the per-ray feature comparison
        import cips3d_b200
        style = G.mapping_network(zs_d["z_nerf"], zs_d["z_inr"])
        fea = cips3d_b200.ops.render_features(
            G.siren.kernel_weights(), G.siren.kernel_film(style), r["c2w"].to(dev), draws["jitter_u"].to(dev),
            draws["pdf_u"].to(dev), None, None, img_size=R, fov=kw["fov"], ray_start=kw["ray_start"],
            ray_end=kw["ray_end"], num_steps=S, hierarchical_sample=True, clamp_mode="relu", noise_std=0.0)["pixels_fea"]
    if str(dev).startswith("cuda"):
        torch.cuda.synchronize()
    # sigma of the last sorted sample, from the oracle
    sig = torch.cat([r["fine"][..., 32], r["coarse"][..., 32]], -1)          # (B,N,2S) in cat order
    z = torch.cat([r["fine_z"], r["z"]], -1)
    sig_last = torch.gather(sig, -1, z.argmax(-1, keepdim=True))[..., 0]       # (B,N)
    step_mask = sig_last.abs() < SIGMA_EPS * sig.abs().max()

    def per_ray(a, b):
        a, b = a.double().cpu(), b.double()
        return (a - b).abs().amax(-1) / b.abs().max().clamp_min(1e-30)

    e_img = per_ray(img.permute(0, 2, 3, 1).reshape(B, R * R, 3), ref_img.permute(0, 2, 3, 1).reshape(B, R * R, 3))
    e_fea = per_ray(fea, r["pixels_fea"])
    out = dict(R=R, B=B, sigma_bias=sigma_bias, rays=B * R * R, step_mask_frac=step_mask.double().mean().item(),
               pitch_yaw_ok=bool(torch.allclose(py.cpu(), ref_py, atol=1e-5)))
    for name, e, a, b in (("img", e_img, img, ref_img), ("fea", e_fea, fea, r["pixels_fea"])):
        bad = e > TOL
        out[name] = dict(max_rel=e.max().item(), l2_rel=rel_err(a.cpu(), b)[1],
                         outlier_frac=bad.double().mean().item(), outliers=int(bad.sum()),
                         outliers_outside_step_mask=int((bad & ~step_mask).sum()),
                         max_rel_outside_step_mask=(e[~step_mask].max().item() if (~step_mask).any() else 0.0))
    return out


@pytest.mark.parametrize("sigma_bias", [0.0, 0.3])
@pytest.mark.parametrize("R,B", [(64, 4), (128, 2), (256, 1)])
def test_generator_forward_matches_oracle_at_baseline_sizes(pkg, R, B, sigma_bias):
    st = baseline_case(R, B, sigma_bias)
    print("PARITY", json.dumps(st))
    os.makedirs("gpurun_out", exist_ok=True)
    with open("gpurun_out/r02_parity_baseline_sizes.jsonl", "a") as f:
        f.write(json.dumps(st) + "\n")
    assert st["pitch_yaw_ok"]
    for k in ("fea", "img"):
        s = st[k]
        assert s["outliers_outside_step_mask"] == 0, (k, s, "a ray off the reference's own discontinuity exceeds 1e-3")
        assert s["max_rel_outside_step_mask"] <= TOL, (k, s)
        assert s["l2_rel"] < 5e-4, (k, s)
        # step-function rays are at most the mask itself
        assert s["outlier_frac"] <= st["step_mask_frac"] + 1e-12, (k, s, st["step_mask_frac"])


E_HID_REF = {}


@pytest.mark.parametrize("scale", [1.0, 64.0, 1024.0])
def test_cips_fp16_operand_range(pkg, scale):
    """The CIPS chain runs on fp16 tcgen05 operands (the reference: fp32).  The modulated layers have no bias and
    LeakyReLU is positively homogeneous, so the hidden state scales linearly with the per-pixel feature input:
    drive it to ~2e3 (fp16 max 65504) and require
      * the SAME relative accuracy of the hidden state against the fp64 oracle at every scale (no overflow, underflow or
        subnormal loss anywhere in the 18-layer chain: measured 6.93e-4 at scale 1, 64 and 1024, identical to the digit);
      * the RGB error after tanh within 1e-3 at the init scale, and within the stated contract above it: tanh is not
        homogeneous, so the absolute error of the pre-tanh sum grows with |hidden| -- measured 1.4e-5 * max|hidden|
        (profiles/r02a_first_run.md); bound 3e-5 * max|hidden| (1e-3 holds while max|hidden| < ~50)."""
    sd = O.synthetic_state_dict(O.generator_template(), seed=31)
    G = build_generator(DEV, sd)
    B, N = 2, 512
    g = torch.Generator().manual_seed(99)
    x = torch.randn(B, N, 32, generator=g) * scale
    w = torch.randn(B, 512, generator=g)
    with torch.no_grad():
        ref64, hid64 = O.cips_net({k: v.double() for k, v in sd.items()}, x.double(), w.double(), return_hidden=True)
        style = {k: w.to(DEV) for k in G.inr_net.style_dim_dict}
        ws, s1p, dm, rw, rb = G.inr_net.kernel_inputs(style, 9)
        rgb, hid = pkg.ops.cips_forward(x.to(DEV), ws, s1p, dm, rw, rb, impl=pkg._lib.IMPL_TC, return_hidden=True)
    torch.cuda.synchronize()
    e_hid = rel_err(hid.cpu(), hid64.float())[0]
    e_rgb = rel_err(rgb.cpu(), ref64.float())[0]
    hmax = hid64.abs().max().item()
    print("RANGE", json.dumps(dict(scale=scale, hidden_absmax=hmax, e_hid=e_hid, e_rgb=e_rgb,
                                   finite=bool(torch.isfinite(hid).all()))))
    assert torch.isfinite(hid).all() and torch.isfinite(rgb).all()
    assert e_hid < 1e-3, (scale, e_hid)
    E_HID_REF.setdefault("v", e_hid)
    assert abs(e_hid - E_HID_REF["v"]) < 0.05 * E_HID_REF["v"], ("hidden error is not scale-free", scale, e_hid, E_HID_REF)
    assert e_rgb < max(1e-3, 3e-5 * hmax), (scale, e_rgb, hmax)
