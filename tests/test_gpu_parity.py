"""GPU parity tests: the CUDA path (through the C-ABI) against
  (a) the golden outputs of the REAL reference (tests/golden, made by tools/make_golden.py),
  (b) the CPU oracle in fp32 / fp64 on the same seeded inputs,
  (c) size-independent properties at BASELINE.json sizes (ray-subset consistency, partition of
      unity of the compositing weights, batch equivariance, SIMT-vs-tcgen05 agreement).

Tolerances (north_star: 1e-3 relative fp32 on rendered RGB):
  * pre-integration sigma/feature tensors: max-rel <= 2e-4
  * integrated features / images: >= 99.5 % of rays within 1e-3 of the feature scale; the rest are
    rays whose last sigma hovers at 0, where the reference itself is discontinuous
    (delta_last = 1e10, pigan_utils.py:243).
"""
import json
import os

import numpy as np
import pytest
import torch

from oracle import cips3d_oracle as O
from _util import (GEN_CASES, GOLDEN, build_generator, close_frac, draws_sequence, load_gen_case, rel_err,
                   replay_draws)

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.fixture(scope="module")
def pkg():
    import __graft_entry__ as g
    g.build()
    import cips3d_b200
    assert cips3d_b200._lib.load().c3d_device_supported(0) == 1, "tests need an sm_100 device"
    return cips3d_b200


def impls(pkg):
    return [("simt", pkg._lib.IMPL_SIMT), ("tc", pkg._lib.IMPL_TC)]


# ------------------------------------------------------------------ tcgen05 building block
@pytest.mark.parametrize("n,k", [(16, 16), (32, 64), (128, 128), (80, 128), (256, 256), (32, 64)])
@pytest.mark.parametrize("a_in_tmem", [False, True])
def test_umma_selftest(pkg, n, k, a_in_tmem):
    g = torch.Generator().manual_seed(n * 1000 + k)
    a = torch.randn(128, k, generator=g).to(DEV)
    b = torch.randn(n, k, generator=g).to(DEV)
    d = pkg.ops.selftest_umma(a, b, a_in_tmem=a_in_tmem)
    ref = a.half().double() @ b.half().double().T
    err = (d.double() - ref).abs().max().item()
    assert err < 1e-3 * ref.abs().max().item(), (n, k, a_in_tmem, err)


# ------------------------------------------------------------------ renderer vs reference goldens
def _render_case(pkg, name, impl, dtype_oracle=None, debug=True):
    sd, zs, draws, kw, meta, ref = load_gen_case(name)
    G = build_generator(DEV, sd)
    G.impl = impl
    B, R, S = meta["B"], meta["img_size"], kw["num_steps"]
    hier = kw["hierarchical_sample"]
    with torch.no_grad():
        style = G.mapping_network(zs["z_nerf"].to(DEV), zs["z_inr"].to(DEV))
        origin, pitch, yaw = O.camera_origin(draws["yaw_n"], draws["pitch_n"], kw["h_stddev"], kw["v_stddev"])
        c2w = O.cam2world(-origin, origin).to(DEV)
        out = pkg.ops.render_features(
            G.siren.kernel_weights(), G.siren.kernel_film(style), c2w, draws["jitter_u"].to(DEV),
            draws["pdf_u"].to(DEV) if hier else None,
            draws["noise_c"].to(DEV) if hier else None, draws["noise_f"].to(DEV),
            img_size=R, fov=kw["fov"], ray_start=kw["ray_start"], ray_end=kw["ray_end"], num_steps=S,
            hierarchical_sample=hier, clamp_mode=kw.get("clamp_mode", "relu"), noise_std=meta["nerf_noise"],
            white_back=kw.get("white_back", False), last_back=kw.get("last_back", False), impl=impl, debug=debug,
            want_depth=True)
    return out, ref, G, style


@pytest.mark.parametrize("name", GEN_CASES)
@pytest.mark.parametrize("impl_name", ["simt", "tc"])
def test_renderer_matches_reference_golden(pkg, name, impl_name):
    impl = dict(impls(pkg))[impl_name]
    out, ref, _, _ = _render_case(pkg, name, impl)
    mr, l2 = rel_err(out["coarse"].cpu(), ref["coarse"])
    assert mr < 2e-4, f"coarse sigma/features max-rel {mr}"
    # fine depths come from an inverse CDF whose bins can be ~1e-5 wide -> conditioning ~1e-4
    assert rel_err(out["all_z"].cpu(), ref["all_z"])[0] < 2e-4
    frac, worst = close_frac(out["pixels_fea"], ref["pixels_fea"], 1e-3)
    assert frac >= 0.995, f"only {frac:.4f} of rays within 1e-3 (worst {worst:.3e})"
    frac_d, _ = close_frac(out["depth"][..., None], ref["depth"][..., None], 1e-3)
    assert frac_d >= 0.995


@pytest.mark.parametrize("name", ["r16_synth", "r16_trained_noise", "r8_softplus_backs", "r8_nohier_s24", "r8_hier_s24"])
@pytest.mark.parametrize("impl_name", ["simt", "tc"])
def test_generator_forward_matches_reference_golden(pkg, name, impl_name):
    """Whole GeneratorNerfINR.forward through the public class surface, draws replayed."""
    impl = dict(impls(pkg))[impl_name]
    sd, zs, draws, kw, meta, ref = load_gen_case(name)
    G = build_generator(DEV, sd)
    G.impl = impl
    prev = os.environ.get("C3D_IMPL")
    os.environ["C3D_IMPL"] = impl_name
    try:
        zs_d = {k: v.to(DEV) for k, v in zs.items()}
        with torch.no_grad(), replay_draws(draws_sequence(draws, kw["hierarchical_sample"]), DEV):
            img, py = G(zs_d, img_size=meta["img_size"], nerf_noise=meta["nerf_noise"], return_aux_img=True, **kw)
    finally:
        os.environ.pop("C3D_IMPL", None) if prev is None else os.environ.__setitem__("C3D_IMPL", prev)
    assert img.shape == ref["img"].shape
    assert torch.allclose(py.cpu(), ref["pitch_yaw"], atol=1e-5)
    a = img.permute(0, 2, 3, 1).reshape(-1, 3)
    b = ref["img"].permute(0, 2, 3, 1).reshape(-1, 3)
    frac, worst = close_frac(a, b, 1e-3)
    assert frac >= 0.99, f"only {frac:.4f} of pixels within 1e-3 (worst {worst:.3e})"


# ------------------------------------------------------------------ CIPS MLP vs oracle
@pytest.mark.parametrize("impl_name", ["simt", "tc"])
@pytest.mark.parametrize("B,N", [(2, 256), (1, 128), (3, 200)])
def test_cips_matches_oracle(pkg, impl_name, B, N):
    impl = dict(impls(pkg))[impl_name]
    sd = O.synthetic_state_dict(O.generator_template(), seed=31)
    G = build_generator(DEV, sd)
    g = torch.Generator().manual_seed(B * 7 + N)
    x = torch.randn(B, N, 32, generator=g)
    w = torch.randn(B, 512, generator=g)
    with torch.no_grad():
        ref64, hid64 = O.cips_net({k: v.double() for k, v in sd.items()}, x.double(), w.double(), return_hidden=True)
        ref32 = O.cips_net(sd, x, w)
        style = {k: w.to(DEV) for k in G.inr_net.style_dim_dict}
        ws, s1p, dm, rw, rb = G.inr_net.kernel_inputs(style, 9)
        rgb, hid = pkg.ops.cips_forward(x.to(DEV), ws, s1p, dm, rw, rb, impl=impl, return_hidden=True)
    e_ref = rel_err(ref32, ref64.float())[0]
    e_hid = rel_err(hid.cpu(), hid64.float())[0]
    e_rgb = rel_err(rgb.cpu(), ref64.float())[0]
    assert e_hid < 1e-3, f"hidden max-rel {e_hid} (fp32 reference itself: {e_ref})"
    assert e_rgb < 1e-3, f"rgb max-rel {e_rgb}"


# Kernel variants that have only run on the CPU emulation (tests/test_emu_cpu.py) are opt-in for ONE round: a protocol error in
# a tcgen05 kernel traps the context (watchdog) and would take the rest of the suite with it.  Every variant that carried this
# mark in round 1 passed its first hardware run (profiles/r02a_first_run.md) and is a normal test now; the mark stays for the next
# never-run kernel.
experimental = pytest.mark.skipif(os.environ.get("C3D_EXPERIMENTAL", "0") != "1",
                                  reason="hardware-unvalidated kernel variant: set C3D_EXPERIMENTAL=1")


@pytest.mark.parametrize("B,N", [(1, 128), (2, 200), (2, 4096)])
def test_cips_backward_chain_and_fused_training(pkg, B, N):
    """c3d_cips_fwd_train + c3d_cips_bwd on the GPU: the fused training path (CIPSNet.train_backend = 'fused') against the
    torch autograd graph of the same module on the same device (see tests/test_emu_cpu.py for the tolerances' rationale)."""
    sd = O.synthetic_state_dict(O.generator_template(), seed=31)
    G = build_generator(DEV, sd).train()
    net = G.inr_net
    g = torch.Generator().manual_seed(B * 7 + N)
    x = torch.randn(B, N, 32, generator=g).to(DEV)
    w = torch.randn(B, 512, generator=g).to(DEV)
    gout = torch.randn(B, N, 3, generator=g).to(DEV)
    style = {k: w for k in net.style_dim_dict}
    xa = x.clone().requires_grad_()
    (net.forward_torch(xa, style) * gout).sum().backward()
    ref = {n: p.grad.clone() for n, p in net.named_parameters() if p.grad is not None}
    net.zero_grad()
    xb = x.clone().requires_grad_()
    (net.forward_fused_train(xb, style) * gout).sum().backward()
    torch.cuda.synchronize()
    assert rel_err(xb.grad.cpu(), xa.grad.cpu())[1] < 8e-2
    for n, p in net.named_parameters():
        if n in ref and float(ref[n].abs().max()) > 0:
            assert rel_err(p.grad.cpu(), ref[n].cpu())[1] < (2e-3 if n.startswith("to_rgbs") else 8e-2), n


@pytest.mark.parametrize("n,k", [(32, 16), (128, 64), (256, 256), (96, 128)])
def test_umma_pair_selftest(pkg, n, k):
    """tcgen05 cta_group::2 in isolation (run this BEFORE the CTA-pair CIPS kernel: it pins the operand partitioning,
    the multicast commit and the cluster-scope barrier hand-off the emulator assumes)."""
    g = torch.Generator().manual_seed(n * 1000 + k)
    a = torch.randn(256, k, generator=g).to(DEV)
    b = torch.randn(n, k, generator=g).to(DEV)
    d = pkg.ops.selftest_umma_pair(a, b)
    ref = a.half().double() @ b.half().double().T
    assert (d.double() - ref).abs().max().item() < 1e-3 * ref.abs().max().item()


@pytest.mark.parametrize("name", GEN_CASES)
def test_renderer_warp_per_ray_math_matches_reference_golden(pkg, name, monkeypatch):
    """C3D_RAY_MATH=warp (warp-per-ray resampling / merge / compositing) against the reference goldens."""
    monkeypatch.setenv("C3D_RAY_MATH", "warp")
    out, ref, _, _ = _render_case(pkg, name, pkg._lib.IMPL_TC)
    assert rel_err(out["coarse"].cpu(), ref["coarse"])[0] < 2e-4
    assert rel_err(out["all_z"].cpu(), ref["all_z"])[0] < 2e-4
    frac, worst = close_frac(out["pixels_fea"], ref["pixels_fea"], 1e-3)
    assert frac >= 0.995, f"only {frac:.4f} of rays within 1e-3 (worst {worst:.3e})"
    assert close_frac(out["depth"][..., None], ref["depth"][..., None], 1e-3)[0] >= 0.995


@pytest.mark.parametrize("name", GEN_CASES)
def test_renderer_fold_math_matches_reference_golden(pkg, name, monkeypatch):
    """C3D_RAY_MATH=fold (sigma head in the layer-1 epilogue, color_layer_linear after compositing) against the
    reference goldens; caller-visible outputs only (the form has no per-point debug outputs)."""
    monkeypatch.setenv("C3D_RAY_MATH", "fold")
    out, ref, _, _ = _render_case(pkg, name, pkg._lib.IMPL_TC, debug=False)
    assert pkg._lib.load().c3d_debug_ray_math_mode() == (0 if name == "r8_hier_s24" else 2)    # 24 + 24 samples: block form only
    frac, worst = close_frac(out["pixels_fea"], ref["pixels_fea"], 1e-3)
    assert frac >= 0.995, f"only {frac:.4f} of rays within 1e-3 (worst {worst:.3e})"
    assert close_frac(out["depth"][..., None], ref["depth"][..., None], 1e-3)[0] >= 0.995


@pytest.mark.parametrize("B,N", [(1, 256), (2, 512), (3, 200), (4, 4096)])
def test_cips_cta_pair_matches_oracle(pkg, B, N, monkeypatch):
    """The cta_group::2 form of the CIPS kernel (the default wherever an image has an even number of tiles) against the fp64
    oracle and, bit for bit, against the single-CTA kernel (C3D_CIPS_PAIR=0; same fp16 operands, same K order -> identical
    accumulators expected; a tolerance of 1e-6 absorbs a different in-tile summation order of the pair's tensor cores).
    Then the image-only call (fp16 residual stream, see csrc/cips_tc.cu ResT) against the oracle and the fp32-residual image."""
    sd = O.synthetic_state_dict(O.generator_template(), seed=31)
    G = build_generator(DEV, sd)
    g = torch.Generator().manual_seed(B * 7 + N)
    x = torch.randn(B, N, 32, generator=g)
    w = torch.randn(B, 512, generator=g)
    with torch.no_grad():
        ref64, hid64 = O.cips_net({k: v.double() for k, v in sd.items()}, x.double(), w.double(), return_hidden=True)
        style = {k: w.to(DEV) for k in G.inr_net.style_dim_dict}
        ws, s1p, dm, rw, rb = G.inr_net.kernel_inputs(style, 9)
        monkeypatch.setenv("C3D_CIPS_PAIR", "0")
        rgb1, hid1 = pkg.ops.cips_forward(x.to(DEV), ws, s1p, dm, rw, rb, impl=pkg._lib.IMPL_TC, return_hidden=True)
        rgb1i = pkg.ops.cips_forward(x.to(DEV), ws, s1p, dm, rw, rb, impl=pkg._lib.IMPL_TC)
        monkeypatch.setenv("C3D_CIPS_PAIR", "1")
        rgb2, hid2 = pkg.ops.cips_forward(x.to(DEV), ws, s1p, dm, rw, rb, impl=pkg._lib.IMPL_TC, return_hidden=True)
        rgb2i = pkg.ops.cips_forward(x.to(DEV), ws, s1p, dm, rw, rb, impl=pkg._lib.IMPL_TC)
        monkeypatch.setenv("C3D_CIPS_RES16", "0")
        rgb2f = pkg.ops.cips_forward(x.to(DEV), ws, s1p, dm, rw, rb, impl=pkg._lib.IMPL_TC)
        torch.cuda.synchronize()
    assert rel_err(hid2.cpu(), hid64.float())[0] < 1e-3
    assert rel_err(rgb2.cpu(), ref64.float())[0] < 1e-3
    assert rel_err(hid2.cpu(), hid1.cpu())[0] < 1e-6 and rel_err(rgb2.cpu(), rgb1.cpu())[0] < 1e-6
    # image-only calls: fp16 residual stream.  Same bound against the oracle; identical between the two kernels; and with
    # C3D_CIPS_RES16=0 the image-only call is the fp32-residual image bit for bit.
    assert rel_err(rgb2i.cpu(), ref64.float())[0] < 1e-3 and rel_err(rgb1i.cpu(), ref64.float())[0] < 1e-3
    assert rel_err(rgb2i.cpu(), rgb1i.cpu())[0] < 1e-6
    assert torch.equal(rgb2f.cpu(), rgb2.cpu())


@pytest.mark.parametrize("img_size,nb", [(32, 4), (64, 5), (256, 7), (1024, 9)])   # < 32: reference tanh(int 0) raises
def test_cips_early_stop_blocks(pkg, img_size, nb):
    sd = O.synthetic_state_dict(O.generator_template(), seed=32)
    G = build_generator(DEV, sd)
    x = torch.randn(1, 128, 32)
    w = torch.randn(1, 512)
    with torch.no_grad():
        ref = O.cips_net(sd, x, w, img_size=img_size)
        out = G.inr_net(x.to(DEV), {k: w.to(DEV) for k in G.inr_net.style_dim_dict}, img_size=img_size)
    assert G.inr_net._n_blocks(img_size) == nb
    assert rel_err(out.cpu(), ref)[0] < 2e-3 or ref.abs().max() == 0


# ------------------------------------------------------------------ properties at benchmark sizes
@pytest.mark.parametrize("impl_name", ["simt", "tc"])
def test_renderer_properties_r64(pkg, impl_name):
    """FFHQ r64 (BASELINE config 2 geometry): ray-subset consistency (bit-exact), partition of unity
    with last_back, batch permutation equivariance, weights in [0,1]."""
    impl = dict(impls(pkg))[impl_name]
    sd = O.synthetic_state_dict(O.generator_template(), seed=5, sigma_bias=0.3)
    G = build_generator(DEV, sd)
    B, R, S = 3, 64, 12
    HW = R * R
    g = torch.Generator(device=DEV).manual_seed(1)
    kw = O.G_KWARGS
    with torch.no_grad():
        style = G.mapping_network(torch.randn(B, 256, device=DEV, generator=g), torch.randn(B, 512, device=DEV, generator=g))
        film, wts = G.siren.kernel_film(style), G.siren.kernel_weights()
        o, _, _ = O.camera_origin(torch.randn(B, 1), torch.randn(B, 1), 0.3, 0.155)
        c2w = O.cam2world(-o, o).to(DEV)
        ju = torch.rand(B, HW, S, device=DEV, generator=g)
        pu = torch.rand(B * HW, S, device=DEV, generator=g)
        nc = torch.randn(B, HW, S, device=DEV, generator=g)
        nf = torch.randn(B, HW, 2 * S, device=DEV, generator=g)
        args = dict(img_size=R, fov=kw["fov"], ray_start=kw["ray_start"], ray_end=kw["ray_end"], num_steps=S,
                    hierarchical_sample=True, noise_std=0.5, impl=impl)
        full = pkg.ops.render_features(wts, film, c2w, ju, pu, nc, nf, last_back=True, want_weights=True,
                                       want_depth=True, **args)
        w = full["weights"]
        assert (w >= -1e-6).all() and torch.allclose(w.sum(-1), torch.ones_like(w.sum(-1)), atol=1e-5)
        assert torch.isfinite(full["pixels_fea"]).all()
        # subset of rays through ray_idx == same rows of the full render, bit for bit
        idx = torch.randperm(HW, device=DEV, generator=g)[:1000]
        sub = pkg.ops.render_features(wts, film, c2w, ju, pu.view(B, HW, S)[:, idx].reshape(-1, S), nc[:, idx],
                                      nf[:, idx], last_back=True, ray_idx=idx, **args)
        assert torch.equal(sub["pixels_fea"], full["pixels_fea"][:, idx])
        # contiguous chunk via ray_offset (the reference's forward_points chunking)
        off, n = 1024, 512
        chunk = pkg.ops.render_features(wts, film, c2w, ju, pu.view(B, HW, S)[:, off:off + n].reshape(-1, S),
                                        nc[:, off:off + n], nf[:, off:off + n], last_back=True, ray_offset=off,
                                        n_rays=n, **args)
        assert torch.equal(chunk["pixels_fea"], full["pixels_fea"][:, off:off + n])
        # batch permutation equivariance
        perm = torch.tensor([2, 0, 1], device=DEV)
        film_p = {k: v[perm] for k, v in film.items()}
        pp = pkg.ops.render_features(wts, film_p, c2w[perm], ju[perm], pu.view(B, HW, S)[perm].reshape(-1, S),
                                     nc[perm], nf[perm], last_back=True, **args)
        assert torch.equal(pp["pixels_fea"], full["pixels_fea"][perm])


def test_tc_agrees_with_simt_at_r64(pkg):
    """tcgen05 (split-fp16) renderer + CIPS vs the fp32-FMA cross-check at FFHQ r64, B=4."""
    sd = O.synthetic_state_dict(O.generator_template(), seed=6, sigma_bias=0.3)
    G = build_generator(DEV, sd)
    B, R = 4, 64
    kw = dict(O.G_KWARGS)
    zs = {"z_nerf": torch.randn(B, 256, device=DEV), "z_inr": torch.randn(B, 512, device=DEV)}
    imgs = {}
    prev = os.environ.get("C3D_IMPL")
    for name, impl in impls(pkg):
        G.impl = impl
        os.environ["C3D_IMPL"] = name
        torch.manual_seed(123)
        with torch.no_grad():
            imgs[name], _ = G(zs, img_size=R, nerf_noise=0.0, **kw)
    os.environ.pop("C3D_IMPL", None) if prev is None else os.environ.__setitem__("C3D_IMPL", prev)
    a = imgs["tc"].permute(0, 2, 3, 1).reshape(-1, 3)
    b = imgs["simt"].permute(0, 2, 3, 1).reshape(-1, 3)
    frac, worst = close_frac(a, b, 1e-3)
    assert frac >= 0.995, (frac, worst)


def test_ray_index_is_row_major(pkg):
    """bit-exact ray indexing: output slot n of image b is pixel (n // R, n % R) (comm_utils.py:392-395).
    A camera looking down -z with sigma forced high makes depth ~ ray length, which is symmetric about the
    image centre and increasing with |x|,|y|; the permutation check uses ray_idx instead."""
    sd = O.synthetic_state_dict(O.generator_template(), seed=8, sigma_bias=0.5)
    G = build_generator(DEV, sd)
    R, S = 16, 12
    with torch.no_grad():
        style = G.mapping_network(torch.randn(1, 256, device=DEV), torch.randn(1, 512, device=DEV))
        film, wts = G.siren.kernel_film(style), G.siren.kernel_weights()
        o, _, _ = O.camera_origin(torch.zeros(1, 1), torch.zeros(1, 1), 0.3, 0.155)
        c2w = O.cam2world(-o, o).to(DEV)
        ju = torch.rand(1, R * R, S, device=DEV)
        args = dict(img_size=R, fov=12, ray_start=0.88, ray_end=1.12, num_steps=S, hierarchical_sample=False,
                    impl=pkg._lib.IMPL_SIMT)
        full = pkg.ops.render_features(wts, film, c2w, ju, **args)["pixels_fea"]
        for (h, w_) in [(0, 0), (3, 11), (15, 15), (7, 8)]:
            one = pkg.ops.render_features(wts, film, c2w, ju, ray_idx=torch.tensor([h * R + w_], device=DEV),
                                          **args)["pixels_fea"]
            assert torch.equal(one[0, 0], full[0, h * R + w_])
    # the oracle's geometry for the same pixel agrees with the kernel's
    sd_c = sd
    r = O.render_features(sd_c, style["nerf_w0"].cpu(), c2w.cpu(), ju.cpu(), None, img_size=R, fov=12,
                          ray_start=0.88, ray_end=1.12, num_steps=S, hierarchical_sample=False)
    frac, worst = close_frac(full, r["pixels_fea"], 1e-3)
    assert frac >= 0.99, (frac, worst)


# ------------------------------------------------------------------ discriminator ops
@pytest.mark.parametrize("shape", [(3, 5, 8, 8), (2, 7), (4, 6, 5, 3), (1, 128, 64, 64), (2, 3, 1, 1)])
def test_bias_act_matches_oracle(pkg, shape):
    g = torch.Generator().manual_seed(len(shape))
    x = torch.randn(*shape, generator=g)
    b = torch.randn(shape[1], generator=g)
    y = pkg.ops.bias_act(x.to(DEV), b.to(DEV))
    assert torch.equal(y.cpu(), O.bias_act(x, b))                       # bit-exact (one mul, one add, one select)
    gr = torch.randn(*shape, generator=g)
    gi = pkg.ops.bias_act(gr.to(DEV), None, y, act=3, grad=1)
    assert torch.equal(gi.cpu(), O.bias_act(gr, None, y.cpu(), act=3, grad=1))
    assert torch.equal(pkg.ops.bias_act(x.to(DEV), b.to(DEV), act=1).cpu(), O.bias_act(x, b, act=1))
    assert pkg.ops.bias_act(torch.empty(0, 4, device=DEV), b[:4].to(DEV)).numel() == 0


def test_fused_leaky_relu_autograd_double_backward(pkg):
    x = torch.randn(2, 4, 6, 6, device=DEV, requires_grad=True)
    b = torch.randn(4, device=DEV, requires_grad=True)
    y = pkg.ops.fused_leaky_relu(x, b)
    y_ref = torch.nn.functional.leaky_relu(x + b.view(1, -1, 1, 1), 0.2) * 2 ** 0.5
    assert torch.allclose(y, y_ref)
    gy = torch.randn_like(y)
    gx, gb = torch.autograd.grad(y, (x, b), gy, create_graph=True)
    gx_r, gb_r = torch.autograd.grad(y_ref, (x, b), gy, create_graph=True)
    assert torch.allclose(gx, gx_r, atol=1e-6) and torch.allclose(gb, gb_r, atol=1e-4)
    # R1-style second order: d/d(gy) of sum(gx^2)
    gy2 = gy.clone().requires_grad_(True)
    gx2, = torch.autograd.grad(pkg.ops.fused_leaky_relu(x, b), x, gy2, create_graph=True)
    gg, = torch.autograd.grad(gx2.pow(2).sum(), gy2)
    gx2r, = torch.autograd.grad(torch.nn.functional.leaky_relu(x + b.view(1, -1, 1, 1), 0.2) * 2 ** 0.5, x, gy2, create_graph=True)
    ggr, = torch.autograd.grad(gx2r.pow(2).sum(), gy2)
    assert torch.allclose(gg, ggr, atol=1e-5)


@pytest.mark.parametrize("shape,up,down,pad", [
    ((2, 3, 9, 10), 1, 1, (2, 2)), ((2, 3, 9, 10), 1, 1, (1, 1)), ((1, 4, 33, 65), 1, 1, (2, 2)),
    ((2, 2, 8, 8), 2, 1, (2, 1)), ((2, 2, 16, 16), 1, 2, (1, 1)), ((1, 1, 1, 1), 1, 1, (2, 2)),
    ((2, 130, 4, 4), 1, 1, (2, 2))])
def test_upfirdn2d_matches_oracle(pkg, shape, up, down, pad):
    x = torch.randn(*shape, dtype=torch.float64)
    k = O._blur_kernel(torch.float64)
    ref = O.upfirdn2d(x, k, (up, up), (down, down), (pad[0], pad[1], pad[0], pad[1]))
    y = pkg.ops.upfirdn2d(x.float().to(DEV), k.float().to(DEV), up=up, down=down, pad=pad)
    assert y.shape == ref.shape
    assert (y.cpu().double() - ref).abs().max() < 1e-5


@pytest.mark.parametrize("shape,pad", [((4, 16, 256, 256), (2, 2, 2, 2)), ((4, 16, 256, 256), (1, 1, 1, 1)),
                                       ((2, 64, 128, 128), (2, 2, 2, 2)), ((3, 5, 8, 8), (1, 1, 1, 1)), ((1, 3, 70, 36), (2, 1, 0, 3))])
def test_blur_kernel_forms_agree(pkg, shape, pad, monkeypatch):
    """The three forms of the 4x4 FIR fast path -- C3D_BLUR=stream (default: register-streaming), =tile (round 1), =tma (TMA row
    staging) -- against the oracle and against each other."""
    g = torch.Generator().manual_seed(sum(shape))
    x = torch.randn(*shape, generator=g)
    k1 = torch.tensor([1., 3., 3., 1.])
    k = (k1[None] * k1[:, None]) / 64
    ys = {}
    for impl in ("stream", "tile", "tma"):
        monkeypatch.setenv("C3D_BLUR", impl)
        ys[impl] = pkg.ops._upfirdn2d_raw(x.to(DEV), k.to(DEV), (1, 1), (1, 1), pad)
    torch.cuda.synchronize()
    ref = O.upfirdn2d(x, k, (1, 1), (1, 1), pad)
    for impl, y in ys.items():
        assert (y.cpu() - ref).abs().max().item() < 1e-5, impl
        assert (y - ys["stream"]).abs().max().item() < 1e-6, impl


def test_upfirdn2d_autograd(pkg):
    k = O._blur_kernel(torch.float32).to(DEV)
    x = torch.randn(2, 3, 12, 12, device=DEV, requires_grad=True)
    for pad in ((2, 2), (1, 1)):
        y = pkg.ops.upfirdn2d(x, k, pad=pad)
        y_ref = O.upfirdn2d(x, k, pad=(pad[0], pad[1], pad[0], pad[1]))
        assert torch.allclose(y, y_ref, atol=1e-6)
        gy = torch.randn_like(y, requires_grad=True)
        gx, = torch.autograd.grad(y, x, gy, create_graph=True)
        gx_r, = torch.autograd.grad(y_ref, x, gy, create_graph=True)
        assert torch.allclose(gx, gx_r, atol=1e-5)
        gg, = torch.autograd.grad(gx.pow(2).sum(), gy)
        gg_r, = torch.autograd.grad(gx_r.pow(2).sum(), gy)
        assert torch.allclose(gg, gg_r, atol=1e-4)


@pytest.mark.parametrize("name", ["r32_main", "r64_aux_fade"])
def test_discriminator_matches_reference_golden(pkg, name):
    g = np.load(os.path.join(GOLDEN, f"disc_{name}.npz"))
    with open(os.path.join(GOLDEN, "state_dict_contract.json")) as f:
        tmpl = json.load(f)["discriminator"]
    D = pkg.Discriminator_MultiScale_Aux(diffaug=False, max_size=1024, channel_multiplier=2,
                                         first_downsample=False, stddev_group=0).to(DEV).eval()
    D.load_state_dict(O.synthetic_state_dict(tmpl, seed=int(g["seed"])))
    with torch.no_grad():
        out = D(torch.from_numpy(g["x"]).to(DEV), use_aux_disc=bool(g["aux"]), alpha=float(g["alpha"]))[0]
    assert rel_err(out.cpu(), torch.from_numpy(g["out"]))[0] < 2e-3     # cuDNN conv algorithms may use TF32-free fp32 but other orders


def test_discriminator_r1_penalty_runs(pkg):
    """train.py:387-394: gradient of D logits w.r.t. real images with create_graph, then backward."""
    D = pkg.Discriminator_MultiScale_Aux(diffaug=True, max_size=64, channel_multiplier=2,
                                         first_downsample=False, stddev_group=0).to(DEV)
    real = torch.rand(4, 3, 32, 32, device=DEV).mul(2).sub(1).requires_grad_(True)
    pred = D(real, use_aux_disc=True, alpha=0.7)[0]
    grad, = torch.autograd.grad(pred.sum(), real, create_graph=True)
    (grad.pow(2).reshape(4, -1).sum(1).mean() + torch.nn.functional.softplus(-pred).mean()).backward()
    used = ("conv_in.32.", "conv_in.16.", "convs.32.", "convs.16.", "convs.8.", "final_conv", "space_linear", "out_linear")
    got = [bool(p.grad is not None and torch.isfinite(p.grad).all()) for n, p in D.named_parameters()
           if any(u in n for u in used)]
    assert all(got) and len(got) > 20


# ------------------------------------------------------------------ training graph
def test_training_graph_matches_fused_forward_and_backprops(pkg):
    sd = O.synthetic_state_dict(O.generator_template(), seed=9, sigma_bias=0.3)
    G = build_generator(DEV, sd)
    B, R = 2, 16
    kw = dict(O.G_KWARGS)
    zs = {"z_nerf": torch.randn(B, 256, device=DEV), "z_inr": torch.randn(B, 512, device=DEV)}
    torch.manual_seed(4)
    with torch.no_grad():
        fused, _ = G(zs, img_size=R, nerf_noise=0.3, return_aux_img=True, **kw)
    torch.manual_seed(4)
    G.train()
    graph, _ = G(zs, img_size=R, nerf_noise=0.3, return_aux_img=True, **kw)
    a = fused.permute(0, 2, 3, 1).reshape(-1, 3)
    b = graph.detach().permute(0, 2, 3, 1).reshape(-1, 3)
    frac, worst = close_frac(a, b, 1e-3)
    assert frac >= 0.99, (frac, worst)
    graph.square().mean().backward()
    used = [p for n, p in G.named_parameters() if p.grad is not None]
    assert len(used) > 100 and all(torch.isfinite(p.grad).all() for p in used)
    # grad_points < H*W -> part_grad_forward (generator.py:1536)
    torch.manual_seed(4)
    part, py = G(zs, img_size=R, nerf_noise=0.0, grad_points=64, **kw)
    assert part.shape == (B, 3, R, R) and part.requires_grad


def test_freeze_nerf_generator_and_forward_points(pkg):
    sd = O.synthetic_state_dict(O.generator_template(), seed=10, sigma_bias=0.3)
    G = build_generator(DEV, sd, frozen=True)
    G.train()
    B, R = 2, 16
    kw = dict(O.G_KWARGS)
    zs = {"z_nerf": torch.randn(B, 256, device=DEV), "z_inr": torch.randn(B, 512, device=DEV)}
    img, _ = G(zs, img_size=R, nerf_noise=0.0, **kw)
    img.mean().backward()
    assert all(p.grad is None for p in G.siren.parameters())
    assert any(p.grad is not None for p in G.inr_net.parameters())
    with torch.no_grad():
        a, pa = G(zs, img_size=R, forward_points=64, nerf_noise=0.2, **kw)       # chunked RNG order path
        assert a.shape == (B, 3, R, R) and pa.shape == (B, 2)
        c, _ = G.forward_camera_pos_and_lookup(
            zs, img_size=R, h_mean=1.5, v_mean=1.5, camera_pos=torch.tensor([[0., 0., 1.]] * B, device=DEV),
            camera_lookup=torch.tensor([[0., 0., -1.]] * B, device=DEV), **kw)
        assert torch.isfinite(c).all()
        d, _ = G(zs, img_size=R, **{**kw, 'psi': 0.7})                                   # truncation path
        assert torch.isfinite(d).all()
