"""Pin the oracle against the real reference (only where /root/reference exists)."""
import pytest
import torch

import ref_shim
from oracle import cips3d_oracle as O

pytestmark = [pytest.mark.reference,
              pytest.mark.skipif(not ref_shim.reference_available(), reason="no /root/reference")]


@pytest.fixture(scope="module")
def ref_G():
    torch.manual_seed(1234)
    return ref_shim.build_reference_generator().eval()


@pytest.mark.parametrize("noise,kw", [
    (0.0, {}), (0.7, {}), (0.3, dict(clamp_mode="softplus", last_back=True)),
    (0.0, dict(white_back=True, hierarchical_sample=False)),
])
def test_generator_bitwise_vs_reference(ref_G, noise, kw):
    import ref_capture
    G = ref_G
    sd = {k: v.clone() for k, v in G.state_dict().items()}     # reference-constructor init
    torch.manual_seed(11)
    zs = G.get_zs(2)
    args = dict(ref_shim.G_KWARGS)
    args.update(kw)
    log = []
    with torch.no_grad(), ref_capture.record_draws(log):
        img, py = G(zs, img_size=12, nerf_noise=noise, return_aux_img=True, **args)
    draws = ref_capture.draws_from_log(log, hierarchical=args["hierarchical_sample"])
    draws.setdefault("noise_c", None)
    draws.setdefault("pdf_u", None)
    with torch.no_grad():
        img2, py2 = O.generator_forward(sd, zs, draws, img_size=12, nerf_noise=noise,
                                        return_aux_img=True, **args)
    assert (img - img2).abs().max().item() < 1e-6
    assert torch.equal(py, py2)


def test_draw_order_matches_reference(ref_G):
    """oracle.draw_randoms replays the reference's RNG call sequence (SURVEY.md §7 hard part 4)."""
    import ref_capture
    G = ref_G
    zs = {"z_nerf": torch.zeros(2, 256), "z_inr": torch.zeros(2, 512)}
    log = []
    torch.manual_seed(5)
    with torch.no_grad(), ref_capture.record_draws(log):
        G(zs, img_size=8, nerf_noise=0.0, **ref_shim.G_KWARGS)
    d_ref = ref_capture.draws_from_log(log)
    torch.manual_seed(5)
    d = O.draw_randoms(2, 8, 12)
    for k in d:
        assert torch.equal(d[k], d_ref[k]), k


def test_discriminator_vs_reference():
    torch.manual_seed(3)
    D = ref_shim.build_reference_discriminator().eval()
    sd = {k: v.clone() for k, v in D.state_dict().items()}
    x = torch.randn(4, 3, 32, 32)
    with torch.no_grad():
        a = D(x, use_aux_disc=True, alpha=0.6)[0]
        b = O.discriminator_forward(sd, x, use_aux_disc=True, alpha=0.6)
    assert (a - b).abs().max().item() < 1e-6


def test_optimiser_tail_vs_torch_and_reference_ema():
    """oracle.clip_adam_ema_step == clip_grad_norm_ + torch.optim.Adam.step + the reference's EMA.update, bit for bit."""
    import copy
    ref_shim.install()
    from exp.comm import comm_model_utils
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(19, 33), torch.nn.Tanh(), torch.nn.Linear(33, 7))
    net_ema = copy.deepcopy(net)
    ema = comm_model_utils.EMA(source=net, target=net_ema, decay=0.999, start_itr=2)
    opt = torch.optim.Adam(params=[{'params': net.parameters(), 'initial_lr': 2e-3}], lr=2e-3, betas=(0.0, 0.999),
                           weight_decay=0, foreach=False)
    P = [p.detach().clone() for p in net.parameters()]
    M = [torch.zeros_like(p) for p in P]
    V = [torch.zeros_like(p) for p in P]
    E = [p.clone() for p in P]
    for it in range(5):
        x = torch.randn(8, 19)
        opt.zero_grad()
        net(x).square().mean().mul(1e3).backward()
        grads = [p.grad.detach().clone() for p in net.parameters()]
        n_ref = torch.nn.utils.clip_grad_norm_(net.parameters(), 10.0)
        opt.step()
        ema.update(itr=it, source_dict=net.state_dict())
        n = O.clip_adam_ema_step(P, grads, M, V, E, step=it + 1, lr=2e-3, betas=(0.0, 0.999), max_norm=10.0,
                                 ema_decay=0.999 if it >= 2 else None)
        assert torch.equal(n, n_ref)
        for a, b in zip(P, net.parameters()):
            assert torch.equal(a, b.detach())
        for a, b in zip(E, net_ema.parameters()):
            assert torch.equal(a, b.detach())


def _ref_pigan():
    import importlib.util
    import os
    spec = importlib.util.spec_from_file_location(
        "c3d_make_golden_pigan", os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools", "make_golden_pigan.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m.import_reference_pigan()


@pytest.mark.parametrize("cls_name", ["SPATIALSIRENBASELINE", "TALLSIREN"])
def test_pigan_surface_constructor_state_dict_and_field_vs_reference(cls_name):
    """cips3d_b200.pigan vs the real piGAN_lib classes: same state_dict keys / shapes / order, same constructor
    initialisation bit for bit under one seed (same RNG call order), same mapping network and field outputs."""
    import cips3d_b200
    Gref_mod, _, Sref = _ref_pigan()
    torch.manual_seed(123)
    ref = Gref_mod.ImplicitGenerator3d(getattr(Sref, cls_name), z_dim=256)
    torch.manual_seed(123)
    mine = cips3d_b200.pigan.ImplicitGenerator3d(getattr(cips3d_b200.pigan, cls_name), z_dim=256)
    sr, sm = ref.state_dict(), mine.state_dict()
    assert list(sr.keys()) == list(sm.keys())
    assert {k: tuple(v.shape) for k, v in sr.items()} == {k: tuple(v) for k, v in O.pigan_template().items()}
    for k in sr:
        assert torch.equal(sr[k], sm[k]), k
    z = torch.randn(3, 256)
    pts, dirs = torch.randn(3, 50, 3) * 0.1, torch.nn.functional.normalize(torch.randn(3, 50, 3), dim=-1)
    with torch.no_grad():
        fr_r, ph_r = ref.siren.mapping_network(z)
        fr_m, ph_m = mine.siren.mapping_network(z)
        assert torch.equal(fr_r, fr_m) and torch.equal(ph_r, ph_m)
        o_r = ref.siren.forward_with_frequencies_phase_shifts(pts, fr_r, ph_r, ray_directions=dirs)
        o_m = mine.siren.forward_with_frequencies_phase_shifts(pts, fr_m, ph_m, ray_directions=dirs)
        o_o = O.pigan_siren(sr, pts, dirs, fr_r, ph_r, gridwarp=cls_name == "SPATIALSIRENBASELINE")
    assert (o_r - o_m).abs().max().item() < 1e-6 and (o_r - o_o).abs().max().item() < 1e-6
    # generate_avg_frequencies draws the same 10000 latents
    ref.device = ref.siren.device = "cpu"
    mine.device = mine.siren.device = "cpu"
    torch.manual_seed(7)
    a_r = ref.generate_avg_frequencies()
    torch.manual_seed(7)
    a_m = mine.generate_avg_frequencies()
    assert torch.equal(a_r[0], a_m[0]) and torch.equal(a_r[1], a_m[1])


@pytest.mark.parametrize("clamp,last_back,white_back,noise_std,T,Cn", [
    ("relu", True, False, 0.0, 24, 32), ("softplus", False, True, 0.5, 24, 32), ("relu", False, False, 0.7, 12, 32),
    ("relu", True, True, 0.3, 24, 3)])
def test_native_fancy_integration_vs_the_real_function(clamp, last_back, white_back, noise_std, T, Cn):
    """ops.fancy_integration (csrc/integrate_ops.cu on the CPU emulation) against the UNMODIFIED exp/pigan/pigan_utils.py
    function: same signature, same RNG draw (identical seed -> identical noise), same three outputs, same gradient."""
    from _emu import emulated
    ref_shim.install()
    from exp.pigan import pigan_utils as ref_utils
    g = torch.Generator().manual_seed(T * 100 + Cn)
    rs = torch.randn(2, 29, T, Cn + 1, generator=g)
    rs[..., Cn] = (rs[..., Cn] + 0.3) * 8
    z = torch.sort(0.88 + 0.24 * torch.rand(2, 29, T, 1, generator=g), -2).values
    d_rgb = torch.randn(2, 29, Cn, generator=g)
    r0 = rs.clone().requires_grad_()
    torch.manual_seed(77)
    rgb0, depth0, w0 = ref_utils.fancy_integration(r0, z, device="cpu", dim_rgb=Cn, noise_std=noise_std, last_back=last_back,
                                                   white_back=white_back, clamp_mode=clamp)
    (g0,) = torch.autograd.grad(rgb0, r0, d_rgb)
    with emulated(async_mode=0) as pkg:
        r1 = rs.clone().requires_grad_()
        torch.manual_seed(77)
        rgb1, depth1, w1 = pkg.ops.fancy_integration(r1, z, device="cpu", dim_rgb=Cn, noise_std=noise_std, last_back=last_back,
                                                     white_back=white_back, clamp_mode=clamp)
        (g1,) = torch.autograd.grad(rgb1, r1, d_rgb)
        with pytest.raises(AssertionError):
            pkg.ops.fancy_integration(r1, z, device="cpu", dim_rgb=Cn, clamp_mode=None)       # pigan_utils.py:252-253
    assert rgb1.shape == rgb0.shape and depth1.shape == depth0.shape and w1.shape == w0.shape
    assert (w1 - w0.detach()).abs().max().item() < 1e-6
    assert (rgb1.detach() - rgb0.detach()).abs().max().item() < 1e-5
    assert (depth1 - depth0.detach()).abs().max().item() < 1e-5
    assert (g1 - g0).abs().max().item() < 1e-4 * g0.abs().max().item() + 1e-6
    torch.manual_seed(77)
    ref_utils.fancy_integration(rs, z, device="cpu", dim_rgb=Cn, noise_std=noise_std, clamp_mode=clamp)
    a = torch.rand(4)
    with emulated(async_mode=0) as pkg:
        torch.manual_seed(77)
        pkg.ops.fancy_integration(rs, z, device="cpu", dim_rgb=Cn, noise_std=noise_std, clamp_mode=clamp)
        b = torch.rand(4)
    assert torch.equal(a, b)                     # both consumed the same amount of the torch RNG stream


@pytest.mark.parametrize("n,k,det", [(10, 12, False), (10, 12, True), (32, 40, False)])
def test_native_sample_pdf_vs_the_real_function(n, k, det):
    """ops.sample_pdf (emulation): the reference function's signature, its torch.rand / linspace draw and its output."""
    from _emu import emulated
    ref_shim.install()
    from exp.pigan import pigan_utils as ref_utils
    g = torch.Generator().manual_seed(n * 10 + k)
    w = torch.rand(41, n, generator=g) + 1e-5
    edges = torch.sort(0.88 + 0.24 * torch.rand(41, n + 2, generator=g), -1).values
    bins = 0.5 * (edges[:, :-1] + edges[:, 1:])
    torch.manual_seed(3)
    want = ref_utils.sample_pdf(bins, w, k, det=det)
    a = torch.rand(2)
    with emulated(async_mode=0) as pkg:
        torch.manual_seed(3)
        got = pkg.ops.sample_pdf(bins, w, k, det=det)
        b = torch.rand(2)
    assert got.shape == want.shape and (got - want).abs().max().item() < 2e-6
    assert torch.equal(a, b)                      # same RNG consumption (none in det mode)


_PF_HIER = (True, 0.4, dict(clamp_mode="relu", white_back=False, last_back=True))
_PF_FLAT = (False, 0.0, dict(clamp_mode="softplus", white_back=True, last_back=False))


@pytest.mark.parametrize("frozen,backend,case", [(False, "torch", _PF_HIER), (False, "fused", _PF_HIER), (True, "fused", _PF_HIER),
                                                 (False, "fused", _PF_FLAT)])
def test_points_forward_vs_reference(monkeypatch, frozen, backend, case):
    """GeneratorNerfINR[_freeze_NeRF].points_forward (generator.py:1659-1762 / 1972-2078; reference signature) against the
    UNMODIFIED method on identical weights, points and seed -- same three RNG draws, same images, same parameter gradients;
    with the torch ops and with the native integration / resampling / merge ops (CPU emulation)."""
    import cips3d_b200
    from _emu import emulated
    from _util import build_generator
    hier, noise, kw = case
    monkeypatch.setattr(cips3d_b200.generator, "_require_cuda", lambda *a, **k: None)
    torch.manual_seed(21)
    Gr = ref_shim.build_reference_generator(frozen=frozen).train()
    G = build_generator("cpu", {k: v.clone() for k, v in Gr.state_dict().items()}, frozen=frozen).train()
    G.train_integrate = backend
    b, n, s = 2, 37, 12
    g = torch.Generator().manual_seed(5)
    origins = torch.randn(b, n, 3, generator=g) * 0.05 + torch.tensor([0., 0., 1.])
    dirs = torch.nn.functional.normalize(torch.randn(b, n, 3, generator=g) * 0.05 + torch.tensor([0., 0., -1.]), dim=-1)
    z_vals = torch.sort(0.88 + 0.24 * torch.rand(b, n, s, 1, generator=g), -2).values
    points = origins[:, :, None] + dirs[:, :, None] * z_vals
    dirs_exp = dirs[:, :, None].expand(-1, -1, s, -1).contiguous()
    idx = torch.randperm(n, generator=g)[:29]
    zs = {"z_nerf": torch.randn(b, 256, generator=g), "z_inr": torch.randn(b, 512, generator=g)}
    args = dict(transformed_points=points, transformed_ray_directions_expanded=dirs_exp, num_steps=s, hierarchical_sample=hier,
                z_vals=z_vals, nerf_noise=noise, transformed_ray_origins=origins, transformed_ray_directions=dirs,
                return_aux_img=True, idx_grad=idx, **kw)
    out = {}
    for name, model in (("ref", Gr), ("new", G)):
        model.zero_grad()
        with emulated(async_mode=0):
            style = model.mapping_network(**zs)
            torch.manual_seed(99)
            inr, aux = model.points_forward(style_dict=style, **args)
            after = torch.rand(3)
            (inr.square().mean() + (aux.square().mean() if aux.requires_grad else 0)).backward()
        out[name] = (inr.detach(), aux.detach(), after, {k: p.grad.clone() for k, p in model.named_parameters() if p.grad is not None})
    assert torch.equal(out["new"][2], out["ref"][2])                        # same RNG consumption
    assert (out["new"][0] - out["ref"][0]).abs().max().item() < 2e-5 and (out["new"][1] - out["ref"][1]).abs().max().item() < 2e-5
    assert out["new"][3].keys() == out["ref"][3].keys()
    assert any(k.startswith("siren.") for k in out["ref"][3]) == (not frozen)
    for k, gr in out["ref"][3].items():
        if backend == "torch":                  # identical arithmetic to the reference's
            assert (out["new"][3][k] - gr).abs().max().item() < 1e-3 * gr.abs().max().item() + 1e-7, k
        else:   # ~1e-6 differences of pixels_fea flip LeakyReLU gates of |z| ~ 0 units in the CIPS MLP (29 pixels): compare in L2
            assert (out["new"][3][k] - gr).norm().item() < 1e-2 * gr.norm().item() + 1e-7, k


@pytest.mark.parametrize("lock,cam", [(False, False), (True, False), (False, True)])
def test_get_world_points_and_direction_vs_reference(lock, cam):
    """cips3d_b200.comm_utils.get_world_points_and_direction against exp/comm/comm_utils.py:682-763: same signature, the same
    seven outputs, the same RNG order (jitter, then the camera draws), also with a given camera and lock_view_dependence."""
    import cips3d_b200
    ref_shim.install()
    from exp.comm import comm_utils as ref_cu
    kw = dict(batch_size=2, num_steps=12, img_size=9, fov=12, ray_start=0.88, ray_end=1.12, h_stddev=0.3, v_stddev=0.155,
              h_mean=1.5707963, v_mean=1.5707963, sample_dist="gaussian", lock_view_dependence=lock, device="cpu")
    if cam:
        kw.update(camera_pos=torch.tensor([[0.1, 0.2, 0.97], [-0.2, 0.0, 0.98]]), camera_lookup=torch.tensor([[-0.1, -0.2, -0.97], [0.2, 0.0, -0.98]]))
    torch.manual_seed(8)
    want = ref_cu.get_world_points_and_direction(**kw)
    a = torch.rand(3)
    torch.manual_seed(8)
    got = cips3d_b200.comm_utils.get_world_points_and_direction(**kw)
    b = torch.rand(3)
    assert torch.equal(a, b) and len(got) == len(want) == 7
    for i, (g_, w_) in enumerate(zip(got, want)):
        assert g_.shape == w_.shape, i
        assert (g_ - w_).abs().max().item() < 2e-6, i
    assert torch.equal(cips3d_b200.comm_utils.gather_points(got[2], torch.tensor([3, 1])), ref_cu.gather_points(got[2], torch.tensor([3, 1])))


def test_host_helpers_of_the_inference_scripts_vs_reference():
    """comm_utils camera trajectories (bit for bit) and inr_layer_swapping (same parameters touched, same blend)."""
    import copy
    import numpy as np
    import cips3d_b200
    from _util import build_generator
    ref_shim.install()
    from exp.comm import comm_utils as ref_cu
    cu = cips3d_b200.comm_utils
    for a, b in ((cu.get_circle_camera_pos_and_lookup(r=1.1, alpha=0.4, num_samples=7, periods=2),
                  ref_cu.get_circle_camera_pos_and_lookup(r=1.1, alpha=0.4, num_samples=7, periods=2)),
                 (cu.get_circle_camera_pos_and_lookup(), ref_cu.get_circle_camera_pos_and_lookup()),
                 (cu.get_yaw_camera_pos_and_lookup(r=1, num_samples=9), ref_cu.get_yaw_camera_pos_and_lookup(r=1, num_samples=9))):
        for x, y in zip(a, b):
            assert x.dtype == y.dtype and np.array_equal(x, y)
    assert cu.get_yaw_pitch_by_xyz(0.3, -0.2, 0.9) == ref_cu.get_yaw_pitch_by_xyz(0.3, -0.2, 0.9)
    torch.manual_seed(2)
    A = build_generator("cpu").inr_net
    B = copy.deepcopy(A)
    T = build_generator("cpu").inr_net                      # different random init
    cu.inr_layer_swapping(A, T, 0.3, ["64", "1024"], verbose=False)
    ref_cu.inr_layer_swapping(B, T, 0.3, ["64", "1024"], verbose=False)
    for (k, p), q in zip(A.state_dict().items(), B.state_dict().values()):
        assert torch.equal(p, q), k


def test_camera_space_ray_functions_vs_reference():
    """comm_utils.get_initial_rays_trig / perturb_points: reference signatures, bit-identical outputs and RNG consumption."""
    import cips3d_b200
    ref_shim.install()
    from exp.comm import comm_utils as ref_cu
    cu = cips3d_b200.comm_utils
    for res in ((7, 7), (6, 9)):
        a = cu.get_initial_rays_trig(bs=2, num_steps=12, fov=12, resolution=res, ray_start=0.88, ray_end=1.12, device="cpu")
        b = ref_cu.get_initial_rays_trig(bs=2, num_steps=12, fov=12, resolution=res, ray_start=0.88, ray_end=1.12, device="cpu")
        for x, y in zip(a, b):
            assert x.shape == y.shape and torch.equal(x, y)
    torch.manual_seed(4)
    p1, z1 = cu.perturb_points(a[0], a[1], a[2], "cpu")
    r1 = torch.rand(2)
    torch.manual_seed(4)
    p2, z2 = ref_cu.perturb_points(b[0], b[1], b[2], "cpu")
    r2 = torch.rand(2)
    assert torch.equal(p1, p2) and torch.equal(z1, z2) and torch.equal(r1, r2)


@pytest.mark.parametrize("cam", [False, True])
def test_transform_sampled_points_vs_reference(cam):
    import cips3d_b200
    ref_shim.install()
    from exp.comm import comm_utils as ref_cu
    cu = cips3d_b200.comm_utils
    pts, z, d = ref_cu.get_initial_rays_trig(bs=2, num_steps=12, fov=12, resolution=(5, 5), ray_start=0.88, ray_end=1.12, device="cpu")
    kw = dict(h_stddev=0.3, v_stddev=0.155, h_mean=1.5707963, v_mean=1.5707963, mode="gaussian", device="cpu")
    if cam:
        kw.update(camera_pos=torch.tensor([[0.1, 0.2, 0.97], [-0.2, 0.0, 0.98]]), camera_lookup=torch.tensor([[-0.1, -0.2, -0.97], [0.2, 0.0, -0.98]]))
    torch.manual_seed(6)
    want = ref_cu.transform_sampled_points(pts, z, d, **kw)
    r1 = torch.rand(2)
    torch.manual_seed(6)
    got = cu.transform_sampled_points(pts, z, d, **kw)
    r2 = torch.rand(2)
    assert torch.equal(r1, r2)
    for i, (g_, w_) in enumerate(zip(got, want)):
        assert g_.shape == w_.shape and (g_ - w_).abs().max().item() < 1e-6, i


def test_pigan_lib_function_surface_vs_reference():
    """cips3d_b200.pigan.fancy_integration / sample_pdf against piGAN_lib/generators/volumetric_rendering.py (emulation)."""
    import cips3d_b200
    import make_golden_pigan
    from _emu import emulated
    _, VR, _ = make_golden_pigan.import_reference_pigan()
    g = torch.Generator().manual_seed(12)
    rs = torch.randn(2, 21, 24, 4, generator=g)
    rs[..., 3] = (rs[..., 3] + 0.3) * 8
    z = torch.sort(0.88 + 0.24 * torch.rand(2, 21, 24, 1, generator=g), -2).values
    torch.manual_seed(1)
    want = VR.fancy_integration(rs, z, device="cpu", noise_std=0.4, last_back=True, white_back=True, clamp_mode="relu")
    w = torch.rand(19, 10, generator=g) + 1e-5
    bins = torch.sort(0.88 + 0.24 * torch.rand(19, 11, generator=g), -1).values
    want_pdf = VR.sample_pdf(bins, w, 12, det=False)
    with emulated(async_mode=0):
        torch.manual_seed(1)
        got = cips3d_b200.pigan.fancy_integration(rs, z, device="cpu", noise_std=0.4, last_back=True, white_back=True, clamp_mode="relu")
        got_pdf = cips3d_b200.pigan.sample_pdf(bins, w, 12, det=False)
        with pytest.raises(TypeError):
            cips3d_b200.pigan.fancy_integration(rs, z, device="cpu")
    with pytest.raises(TypeError):
        VR.fancy_integration(rs, z, device="cpu")
    for a, b in zip(got, want):
        assert a.shape == b.shape and (a - b).abs().max().item() < 1e-5
    assert (got_pdf - want_pdf).abs().max().item() < 2e-6
