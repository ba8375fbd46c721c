"""GPU parity of the pi-GAN surface (cips3d_b200.pigan on c3d_pigan_render_fwd) against the goldens of the REAL piGAN_lib
classes, through the public class surface with the random draws replayed in the reference's order.  File order: runs after
the CIPS-3D parity suites."""
import os

import pytest
import torch

from _util import PIGAN_CASES, close_frac, load_pigan_case, rel_err, replay_draws

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.fixture(scope="module")
def pkg():
    import __graft_entry__ as g
    g.build()
    import cips3d_b200
    return cips3d_b200


def _build(pkg, sd, meta):
    cls = pkg.pigan.SPATIALSIRENBASELINE if meta["siren_cls"] == "SPATIALSIRENBASELINE" else pkg.pigan.TALLSIREN
    G = pkg.pigan.ImplicitGenerator3d(cls, z_dim=256).to(DEV).eval()
    G.load_state_dict(sd)
    G.device = G.siren.device = DEV
    return G


def _draw_seq(draws, hier):
    seq = [draws["jitter_u"][..., None], draws["yaw_n"], draws["pitch_n"]]
    if hier:
        seq += [draws["noise_c"][..., None], draws["pdf_u"]]
    return seq + [draws["noise_f"][..., None]]


@pytest.mark.parametrize("name", [c for c in PIGAN_CASES if "staged" not in c])
@pytest.mark.parametrize("path", ["native", "torch"])
def test_pigan_forward_matches_reference_golden(pkg, name, path):
    sd, z, draws, kw, meta, ref = load_pigan_case(name)
    G = _build(pkg, sd, meta)
    G.force_torch_path = path == "torch"
    with torch.no_grad(), replay_draws(_draw_seq(draws, kw["hierarchical_sample"]), DEV):
        img, py = G(z.to(DEV), img_size=meta["img_size"], nerf_noise=meta["nerf_noise"], **kw)
    assert torch.allclose(py.cpu(), ref["pitch_yaw"], atol=1e-5)
    a = img.permute(0, 2, 3, 1).reshape(-1, 3)
    b = ref["img"].permute(0, 2, 3, 1).reshape(-1, 3)
    frac, worst = close_frac(a, b, 1e-3)
    assert frac >= 0.99, f"only {frac:.4f} of pixels within 1e-3 (worst {worst:.3e})"


def test_pigan_staged_forward_matches_reference_golden(pkg):
    """staged_forward (truncation psi = 0.7).  The fixture stores the reference's averaged frequencies (the mean over its
    10000 latent draws) instead of the 10 MB of latents, so generate_avg_frequencies is pinned to them here."""
    sd, z, draws, kw, meta, ref = load_pigan_case("spatial_r6_staged_psi07")
    G = _build(pkg, sd, meta)
    G.avg_frequencies, G.avg_phase_shifts = (t.to(DEV) for t in meta["avg"])
    G.generate_avg_frequencies = lambda: (G.avg_frequencies, G.avg_phase_shifts)
    with replay_draws(_draw_seq(draws, kw["hierarchical_sample"]), DEV):
        img, depth = G.staged_forward(z.to(DEV), img_size=meta["img_size"], nerf_noise=meta["nerf_noise"], psi=meta["staged_psi"], **kw)
    assert not img.is_cuda and not depth.is_cuda and depth.shape == (meta["B"], meta["img_size"], meta["img_size"])
    frac, worst = close_frac(img.permute(0, 2, 3, 1).reshape(-1, 3), ref["img"].permute(0, 2, 3, 1).reshape(-1, 3), 1e-3)
    assert frac >= 0.99, (frac, worst)


def test_pigan_training_graph_backprops_and_matches_native(pkg):
    sd, z, draws, kw, meta, ref = load_pigan_case("spatial_r8")
    G = _build(pkg, sd, meta).train()
    hier = kw["hierarchical_sample"]
    with replay_draws(_draw_seq(draws, hier), DEV):
        img, _ = G(z.to(DEV), img_size=meta["img_size"], nerf_noise=0.0, **kw)       # parameters require grad -> torch graph
    img.square().mean().backward()
    assert all(p.grad is not None and torch.isfinite(p.grad).all() for p in G.siren.parameters())
    with torch.no_grad(), replay_draws(_draw_seq(draws, hier), DEV):
        img2, _ = G(z.to(DEV), img_size=meta["img_size"], nerf_noise=0.0, **kw)      # native kernels
    frac, _ = close_frac(img.detach().permute(0, 2, 3, 1).reshape(-1, 3), img2.permute(0, 2, 3, 1).reshape(-1, 3), 1e-3)
    assert frac >= 0.99


@pytest.mark.parametrize("impl,pair", [("tc", "0"), ("tc", "1"), ("simt", "0")])
@pytest.mark.parametrize("name", [c for c in PIGAN_CASES if "staged" not in c])
def test_pigan_renderer_variants_match_reference_golden(pkg, name, impl, pair, monkeypatch):
    """Every native pi-GAN renderer through the class surface: the fused tcgen05 kernel (pigan_tc.cu; the default since its
    first hardware run, profiles/r02a_first_run.md), its CTA-pair form, and the fp32-FMA per-layer cross-check (pigan_simt.cu)."""
    monkeypatch.setenv("C3D_PIGAN_IMPL", impl)
    monkeypatch.setenv("C3D_PIGAN_PAIR", pair)
    sd, z, draws, kw, meta, ref = load_pigan_case(name)
    G = _build(pkg, sd, meta)
    with torch.no_grad(), replay_draws(_draw_seq(draws, kw["hierarchical_sample"]), DEV):
        img, py = G(z.to(DEV), img_size=meta["img_size"], nerf_noise=meta["nerf_noise"], **kw)
    torch.cuda.synchronize()
    frac, worst = close_frac(img.permute(0, 2, 3, 1).reshape(-1, 3), ref["img"].permute(0, 2, 3, 1).reshape(-1, 3), 1e-3)
    assert frac >= 0.99, f"only {frac:.4f} of pixels within 1e-3 (worst {worst:.3e})"
