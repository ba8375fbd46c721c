"""CPU tests of the CUDA kernels themselves, through the functional emulation in tools/emu.

The product sources (csrc/*.cu) are compiled with g++ -DC3D_EMU; every CUDA thread becomes a fiber, and mbarriers,
tensor memory, tcgen05.mma (decoded from the real descriptors), bulk copies, named / cluster barriers are emulated
with asynchronous completion under a seeded schedule (see tools/emu/c3d_emu.h).  The same C-ABI entry points and
the same Python host layer as on the GPU run here, against the same goldens of the real reference and the same
oracle -- so layout, descriptor, index and protocol errors in a kernel show up without a GPU.  The emulator itself
is pinned by the hardware-validated kernels (they must reproduce the goldens) and by a fault-injection case (the
parity-aliasing race of the round-1 weight ring must be reported)."""
import ctypes as C

import pytest
import torch

import _emu
from _emu import emulated
from _util import (GEN_CASES, PIGAN_CASES, build_generator, close_frac, load_gen_case, load_pigan_case, pigan_freq_phase,
                   rel_err)
from oracle import cips3d_oracle as O

TC, SIMT = 0, 1
MODES = {"eager": 0, "lazy": 1, "random": 2}


def test_product_loader_refuses_emulation_build():
    import cips3d_b200
    L = cips3d_b200._lib
    path = _emu.build_emu.build()
    saved = (L._lib, L.LIB_PATH)
    L._lib, L.LIB_PATH = None, path
    try:
        with pytest.raises(L.C3dError, match="emulation"):
            L.load()
    finally:
        L._lib, L.LIB_PATH = saved


def test_emulation_build_exports_the_whole_c_abi():
    import cips3d_b200
    lib = _emu.emu_lib()
    for name in cips3d_b200._lib.EXPORTS:
        assert hasattr(lib, name), name
    assert lib.c3d_emulated() == 1


@pytest.mark.parametrize("mode", ["eager", "lazy", "random"])
@pytest.mark.parametrize("n,k,a_in_tmem", [(16, 16, False), (128, 128, True), (80, 128, True), (256, 256, False), (32, 64, True)])
def test_emu_umma_selftest(mode, n, k, a_in_tmem):
    g = torch.Generator().manual_seed(n * 1000 + k)
    a, b = torch.randn(128, k, generator=g), torch.randn(n, k, generator=g)
    with emulated(async_mode=MODES[mode], seed=n + k) as pkg:
        d = pkg.ops.selftest_umma(a, b, a_in_tmem=a_in_tmem)
    ref = a.half().double() @ b.half().double().T
    assert (d.double() - ref).abs().max().item() < 1e-5 * ref.abs().max().item()


@pytest.mark.parametrize("mode", ["eager", "lazy", "random"])
@pytest.mark.parametrize("n,k", [(32, 16), (128, 64), (256, 256), (96, 128)])
def test_emu_umma_pair_selftest(mode, n, k):
    g = torch.Generator().manual_seed(n * 1000 + k)
    a, b = torch.randn(256, k, generator=g), torch.randn(n, k, generator=g)
    with emulated(async_mode=MODES[mode], seed=n + k) as pkg:
        d = pkg.ops.selftest_umma_pair(a, b)
    ref = a.half().double() @ b.half().double().T
    assert (d.double() - ref).abs().max().item() < 1e-5 * ref.abs().max().item()


def _render(pkg, name, impl, debug=True, **extra):
    sd, zs, draws, kw, meta, ref = load_gen_case(name)
    G = build_generator("cpu", sd)
    R, S, hier = meta["img_size"], kw["num_steps"], kw["hierarchical_sample"]
    with torch.no_grad():
        style = G.mapping_network(zs["z_nerf"], zs["z_inr"])
        origin, _, _ = O.camera_origin(draws["yaw_n"], draws["pitch_n"], kw["h_stddev"], kw["v_stddev"])
        c2w = O.cam2world(-origin, origin)
        out = pkg.ops.render_features(
            G.siren.kernel_weights(), G.siren.kernel_film(style), c2w, draws["jitter_u"],
            draws["pdf_u"] if hier else None, draws["noise_c"] if hier else None, draws["noise_f"],
            img_size=R, fov=kw["fov"], ray_start=kw["ray_start"], ray_end=kw["ray_end"], num_steps=S,
            hierarchical_sample=hier, clamp_mode=kw.get("clamp_mode", "relu"), noise_std=meta["nerf_noise"],
            white_back=kw.get("white_back", False), last_back=kw.get("last_back", False), impl=impl, debug=debug, **extra)
    return out, ref


def _check_render(out, ref):
    # same bounds as tests/test_gpu_parity.py::test_renderer_matches_reference_golden
    assert rel_err(out["coarse"], ref["coarse"])[0] < 2e-4
    assert rel_err(out["all_z"], ref["all_z"])[0] < 2e-4
    frac, worst = close_frac(out["pixels_fea"], ref["pixels_fea"], 1e-3)
    assert frac >= 0.995, (frac, worst)
    assert close_frac(out["depth"][..., None], ref["depth"][..., None], 1e-3)[0] >= 0.995


@pytest.mark.parametrize("mode", ["lazy", "random"])
@pytest.mark.parametrize("name", GEN_CASES)
def test_emu_ray_siren_tc_matches_reference_golden(name, mode):
    """The fused tcgen05 renderer kernel, executed thread by thread on the CPU, against the REAL reference."""
    with emulated(async_mode=MODES[mode], seed=11) as pkg:
        out, ref = _render(pkg, name, TC)
    _check_render(out, ref)


@pytest.mark.parametrize("mode", ["lazy", "random"])
@pytest.mark.parametrize("name", GEN_CASES)
def test_emu_ray_siren_tc_warp_per_ray_math(name, mode, monkeypatch):
    """C3D_RAY_MATH=warp: resampling / merge / compositing by one warp per ray with shuffles (scans, rank sort) instead of
    block-wide shared-memory passes.  Same goldens of the real reference, same bounds (noise, softplus, last_back,
    white_back, non-hierarchical S = 24 are all in the cases).  Not yet timed on hardware."""
    monkeypatch.setenv("C3D_RAY_MATH", "warp")
    with emulated(async_mode=MODES[mode], seed=5) as pkg:
        out, ref = _render(pkg, name, TC)
    _check_render(out, ref)


@pytest.mark.parametrize("mode", ["lazy", "random"])
@pytest.mark.parametrize("name", GEN_CASES)
def test_emu_ray_siren_tc_fold_math(name, mode, monkeypatch):
    """C3D_RAY_MATH=fold: warp-per-ray math, the sigma head as an fp32 dot product in the layer-1 epilogue and
    color_layer_linear applied once per ray to the composited colour sines (3 MMA phases per pass instead of 4).  The
    per-point debug outputs do not exist in this form (the library falls back to warp math when they are requested), so
    the check is on the outputs a caller sees: pixel features, depth, weights.  Not yet timed on hardware."""
    monkeypatch.setenv("C3D_RAY_MATH", "fold")
    with emulated(async_mode=MODES[mode], seed=7) as pkg:
        out, ref = _render(pkg, name, TC, debug=False, want_depth=True, want_weights=True)
        # the fold form ran, not a fallback -- except for 24 + 24 samples per ray, which only the block-wide form serves (2S <= 32)
        assert _emu.emu_lib().c3d_debug_ray_math_mode() == (0 if name == "r8_hier_s24" else 2)
    frac, worst = close_frac(out["pixels_fea"], ref["pixels_fea"], 1e-3)
    assert frac >= 0.995, (frac, worst)
    assert close_frac(out["depth"][..., None], ref["depth"][..., None], 1e-3)[0] >= 0.995


@pytest.mark.parametrize("name", ["r16_trained_noise", "r8_nohier_s24"])
def test_emu_ray_siren_simt_matches_reference_golden(name):
    with emulated(async_mode=0) as pkg:
        out, ref = _render(pkg, name, SIMT)
    _check_render(out, ref)


def _cips_case(pkg, B, N, impl, seed=31, n_blocks=9):
    sd = O.synthetic_state_dict(O.generator_template(), seed=seed)
    G = build_generator("cpu", sd)
    g = torch.Generator().manual_seed(B * 7 + N)
    x, w = torch.randn(B, N, 32, generator=g), torch.randn(B, 512, generator=g)
    with torch.no_grad():
        ref64, hid64 = O.cips_net({k: v.double() for k, v in sd.items()}, x.double(), w.double(), return_hidden=True)
        ws, s1p, dm, rw, rb = G.inr_net.kernel_inputs({k: w for k in G.inr_net.style_dim_dict}, n_blocks)
        rgb, hid = pkg.ops.cips_forward(x, ws, s1p, dm, rw, rb, impl=impl, return_hidden=True)
    return rel_err(hid, hid64.float())[0], rel_err(rgb, ref64.float())[0]


@pytest.mark.parametrize("mode", ["eager", "lazy", "random"])
@pytest.mark.parametrize("B,N", [(1, 128), (3, 200)])
def test_emu_cips_tc_matches_oracle(B, N, mode, monkeypatch):
    """The fused single-CTA CIPS tcgen05 kernel (weight ring, dual issuers, staircase pipeline) on the CPU vs the fp64 oracle;
    (3, 200) has ragged tiles and more tiles than emulated SMs (the persistent loop and the dummy tiles run)."""
    monkeypatch.setenv("C3D_CIPS_PAIR", "0")
    with emulated(async_mode=MODES[mode], seed=B + N, sms=2) as pkg:
        e_hid, e_rgb = _cips_case(pkg, B, N, TC)
    assert e_hid < 1e-3 and e_rgb < 1e-3, (e_hid, e_rgb)


@pytest.mark.parametrize("cl", [2, 4])
def test_emu_cips_tc_cluster_multicast(cl, monkeypatch):
    """C3D_CIPS_CLUSTER=2|4: weight tiles multicast across a thread-block cluster, remote mbarrier arrives."""
    monkeypatch.setenv("C3D_CIPS_CLUSTER", str(cl))
    with emulated(async_mode=2, seed=cl, sms=4) as pkg:
        e_hid, e_rgb = _cips_case(pkg, 2, 128 * cl, TC)
        assert e_hid < 1e-3 and e_rgb < 1e-3, (e_hid, e_rgb)
        # 3 tiles per image: a cluster would straddle two images' weight streams -> the host must fall back to CL = 1
        e_hid, e_rgb = _cips_case(pkg, 2, 384, TC)
        assert e_hid < 1e-3 and e_rgb < 1e-3, (e_hid, e_rgb)


@pytest.mark.parametrize("mode", ["eager", "lazy", "random"])
@pytest.mark.parametrize("B,N,sms", [(1, 256, 2), (2, 512, 2), (3, 256, 4), (2, 200, 2)])
def test_emu_cips_tc_cta_pair(B, N, sms, mode, monkeypatch):
    """C3D_CIPS_PAIR=1: tcgen05 cta_group::2 -- the leader CTA issues M = 256 MMAs for both CTAs of a cluster, each
    CTA streams half of every weight tile, the peer relays its fills and both epilogues report to the leader.
    (2, 512): four iterations of the persistent loop; (3, 256, sms 4): two pairs, one of them with a dummy tail
    iteration; (2, 200): ragged second tile.  Not yet run on hardware (opt-in), the emulator's cta_group::2 model
    follows cute's MMA_Traits<SM100_MMA_F16BF16_2x1SM_SS> operand partitioning."""
    monkeypatch.setenv("C3D_CIPS_PAIR", "1")
    with emulated(async_mode=MODES[mode], seed=B * N + sms, sms=sms) as pkg:
        e_hid, e_rgb = _cips_case(pkg, B, N, TC)
    assert e_hid < 1e-3 and e_rgb < 1e-3, (e_hid, e_rgb)


@pytest.mark.parametrize("pair", ["0", "1"])
@pytest.mark.parametrize("B,N", [(2, 256), (1, 384)])
def test_emu_cips_image_only_call_uses_the_fp16_residual_stream(B, N, pair, monkeypatch):
    """Without a hidden-state output the kernel keeps the skip connections' stream as fp16 (csrc/cips_tc.cu ResT): same bound against
    the fp64 oracle as the fp32 stream; C3D_CIPS_RES16=0 gives back the fp32-stream image bit for bit."""
    monkeypatch.setenv("C3D_CIPS_PAIR", pair)
    sd = O.synthetic_state_dict(O.generator_template(), seed=31)
    G = build_generator("cpu", sd)
    g = torch.Generator().manual_seed(B * 7 + N)
    x, w = torch.randn(B, N, 32, generator=g), torch.randn(B, 512, generator=g)
    with torch.no_grad():
        ref64 = O.cips_net({k: v.double() for k, v in sd.items()}, x.double(), w.double())
        with emulated(async_mode=2, seed=N, sms=2) as pkg:
            ws, s1p, dm, rw, rb = G.inr_net.kernel_inputs({k: w for k in G.inr_net.style_dim_dict}, 9)
            rgb_h, _ = pkg.ops.cips_forward(x, ws, s1p, dm, rw, rb, impl=TC, return_hidden=True)
            rgb16 = pkg.ops.cips_forward(x, ws, s1p, dm, rw, rb, impl=TC)
            monkeypatch.setenv("C3D_CIPS_RES16", "0")
            rgb32 = pkg.ops.cips_forward(x, ws, s1p, dm, rw, rb, impl=TC)
    assert rel_err(rgb16, ref64.float())[0] < 1e-3
    assert torch.equal(rgb32, rgb_h) and not torch.equal(rgb16, rgb_h)


def test_emu_cips_tc_cta_pair_falls_back_on_odd_tile_counts(monkeypatch):
    monkeypatch.setenv("C3D_CIPS_PAIR", "1")
    with emulated(async_mode=2, sms=2) as pkg:
        e_hid, e_rgb = _cips_case(pkg, 2, 384, TC)       # 3 tiles per image: a pair would straddle two images
    assert e_hid < 1e-3 and e_rgb < 1e-3, (e_hid, e_rgb)


def test_emu_cips_simt_matches_oracle():
    with emulated(async_mode=0) as pkg:
        e_hid, e_rgb = _cips_case(pkg, 1, 128, SIMT)
    assert e_hid < 1e-3 and e_rgb < 1e-3, (e_hid, e_rgb)


def test_emu_detects_the_round1_weight_ring_race(monkeypatch):
    """Fault injection: -DC3D_INJECT_RING_RACE restores the round-1 protocol (an MMA issuer skips the fills of the
    tiles it does not own).  On hardware it only failed under ncu's replay timing; the emulator must report it
    whenever asynchronous operations complete late."""
    monkeypatch.setenv("C3D_CIPS_PAIR", "0")       # the two-issuer ring is the single-CTA kernel's
    path = _emu.build_emu.build(extra_defs=("C3D_INJECT_RING_RACE",), tag="ringrace")
    good = _emu._cdll
    bad = C.CDLL(path)
    bad.c3d_emu_configure.argtypes = [C.c_int, C.c_ulonglong, C.c_int, C.c_int]
    _emu._cdll = bad
    try:
        failures = 0
        for seed in (1, 2, 3):
            with emulated(async_mode=MODES["lazy"], seed=seed) as pkg:
                try:
                    e_hid, e_rgb = _cips_case(pkg, 1, 256, TC)
                    failures += int(not (e_hid < 1e-3 and e_rgb < 1e-3))
                except pkg._lib.C3dError:
                    failures += 1
        assert failures == 3, f"the injected race was reported in only {failures} of 3 schedules"
    finally:
        _emu._cdll = good


@pytest.mark.parametrize("shape", [(3, 5, 8, 8), (2, 7), (4, 6, 5, 3), (2, 3, 1, 1)])
def test_emu_bias_act(shape):
    g = torch.Generator().manual_seed(sum(shape))
    x, b = torch.randn(*shape, generator=g), torch.randn(shape[1], generator=g)
    with emulated(async_mode=0) as pkg:
        y = pkg.ops.bias_act(x, b)
        ref = torch.randn(*shape, generator=g)
        yg = pkg.ops.bias_act(x, None, ref, act=3, grad=1)
    assert torch.equal(y, O.bias_act(x, b))
    assert torch.equal(yg, O.bias_act(x, None, ref, act=3, grad=1))


@pytest.mark.parametrize("shape,up,down,pad", [
    ((2, 3, 9, 11), (1, 1), (1, 1), (2, 2, 2, 2)), ((1, 4, 64, 64), (1, 1), (1, 1), (1, 1, 1, 1)),
    ((1, 2, 8, 8), (2, 2), (1, 1), (2, 1, 2, 1)), ((2, 2, 16, 12), (1, 1), (2, 2), (1, 1, 1, 1))])
def test_emu_upfirdn2d(shape, up, down, pad):
    g = torch.Generator().manual_seed(sum(shape))
    x = torch.randn(*shape, generator=g)
    k = torch.tensor([1., 3., 3., 1.])
    k = (k[None] * k[:, None]) / 64
    with emulated(async_mode=0) as pkg:
        y = pkg.ops._upfirdn2d_raw(x, k, up, down, pad)
    ref = O.upfirdn2d(x, k, up, down, pad)
    assert y.shape == ref.shape and (y - ref).abs().max().item() < 1e-5


@pytest.mark.parametrize("mode", ["eager", "lazy", "random"])
@pytest.mark.parametrize("shape,pad", [
    ((2, 3, 64, 64), (2, 2, 2, 2)), ((1, 2, 256, 256), (1, 1, 1, 1)), ((1, 2, 256, 256), (2, 2, 2, 2)),
    ((3, 1, 8, 8), (2, 2, 2, 2)), ((1, 5, 4, 4), (1, 1, 1, 1)), ((1, 1, 100, 36), (2, 1, 0, 3)), ((1, 2, 70, 128), (1, 1, 1, 1))])
def test_emu_blur_tma_streaming_kernel(shape, pad, mode, monkeypatch):
    """C3D_BLUR_TMA=1: the discriminator's 4x4 blur with TMA row staging (cp.async.bulk per input row into zero-margined
    shared rows, two mbarrier-guarded buffers, persistent strips).  2 emulated SMs -> every CTA walks several strips, so
    the prefetch / re-staging protocol runs; shapes cover both paddings D uses at 256^2 (out 257 / 255: main columns +
    tail), planes smaller than one strip, ragged last strips, asymmetric pads, a non-separable random kernel."""
    monkeypatch.setenv("C3D_BLUR_TMA", "1")
    g = torch.Generator().manual_seed(sum(shape))
    x = torch.randn(*shape, generator=g)
    k1 = torch.tensor([1., 3., 3., 1.])
    for k in ((k1[None] * k1[:, None]) / 64, torch.randn(4, 4, generator=g)):
        with emulated(async_mode=MODES[mode], seed=shape[2], sms=2) as pkg:
            y = pkg.ops._upfirdn2d_raw(x, k, (1, 1), (1, 1), pad)
        ref = O.upfirdn2d(x, k, (1, 1), (1, 1), pad)
        assert y.shape == ref.shape and (y - ref).abs().max().item() < 1e-5


@pytest.mark.parametrize("impl", ["stream", "tile"])
@pytest.mark.parametrize("shape,pad", [
    ((2, 3, 64, 64), (2, 2, 2, 2)), ((1, 2, 256, 256), (1, 1, 1, 1)), ((1, 2, 256, 256), (2, 2, 2, 2)),
    ((3, 1, 8, 8), (2, 2, 2, 2)), ((1, 5, 4, 4), (1, 1, 1, 1)), ((1, 1, 100, 36), (2, 1, 0, 3)), ((1, 2, 70, 128), (1, 1, 1, 1)),
    ((1, 1, 33, 300), (1, 2, 2, 1)), ((1, 1, 5, 2), (3, 3, 3, 3)), ((1, 2, 37, 65), (-1, 2, 2, -1))])
def test_emu_blur_register_streaming_kernel(shape, pad, impl, monkeypatch):
    """C3D_BLUR=stream (the default 4x4 FIR path: thread = output column marching down a 32-row strip, the window in
    registers, no shared memory) and the round-1 tile form, against the oracle: both paddings D uses at 256^2 (out 257 / 255),
    planes smaller than a strip, ragged last strips, more than 8 column warps, asymmetric and NEGATIVE pads (crop), a
    non-separable random kernel."""
    monkeypatch.setenv("C3D_BLUR", impl)
    g = torch.Generator().manual_seed(sum(shape))
    x = torch.randn(*shape, generator=g)
    k1 = torch.tensor([1., 3., 3., 1.])
    for k in ((k1[None] * k1[:, None]) / 64, torch.randn(4, 4, generator=g)):
        with emulated(async_mode=0, sms=2) as pkg:
            y = pkg.ops._upfirdn2d_raw(x, k, (1, 1), (1, 1), pad)
        ref = O.upfirdn2d(x, k, (1, 1), (1, 1), pad)
        assert y.shape == ref.shape and (y - ref).abs().max().item() < 1e-5


@pytest.mark.parametrize("impl,mode", [("simt", "eager"), ("tc", "lazy"), ("tc", "random"), ("tc-pair", "lazy"), ("tc-pair", "random")])
@pytest.mark.parametrize("name", PIGAN_CASES)
def test_emu_pigan_renderer_matches_reference_golden(name, impl, mode, monkeypatch):
    """c3d_pigan_render_fwd (rays -> 8 x 256 FiLM-SIREN with view-dependent colour -> resampling -> rgb compositing) against
    the REAL piGAN_lib ImplicitGenerator3d + SPATIALSIRENBASELINE / TALLSIREN: forward, noise + softplus + backs,
    lock_view_dependence, non-hierarchical S = 24, staged_forward's truncated frequencies."""
    # impl "tc": pigan_tc.cu -- one fused tcgen05 kernel (streamed 256x256 hi/lo weight tiles, A operand in TMEM, 3-pass
    # split precision, heads and the view-direction columns in the epilogues, warp-per-ray math); "simt": pigan_simt.cu
    # "tc-pair": the same kernel as a tcgen05 cta_group::2 CTA pair (two ray groups per weight stream, A from each CTA's TMEM)
    monkeypatch.setenv("C3D_PIGAN_IMPL", impl.split("-")[0])
    monkeypatch.setenv("C3D_PIGAN_PAIR", "1" if impl.endswith("pair") else "0")
    sd, z, draws, kw, meta, ref = load_pigan_case(name)
    with emulated(async_mode=MODES[mode], seed=9) as pkg:
        cls = pkg.pigan.SPATIALSIRENBASELINE if meta["siren_cls"] == "SPATIALSIRENBASELINE" else pkg.pigan.TALLSIREN
        G = pkg.pigan.ImplicitGenerator3d(cls, z_dim=256)
        G.load_state_dict(sd)
        hier = kw["hierarchical_sample"]
        with torch.no_grad():
            fr, ph = pigan_freq_phase(sd, z, meta)
            origin, _, _ = O.camera_origin(draws["yaw_n"], draws["pitch_n"], kw["h_stddev"], kw["v_stddev"], kw["h_mean"], kw["v_mean"])
            out = pkg.ops.pigan_render(
                G.siren.kernel_weights(fr, ph), O.cam2world(-origin, origin), draws["jitter_u"], draws["pdf_u"] if hier else None,
                draws["noise_c"] if hier else None, draws["noise_f"], img_size=meta["img_size"], fov=kw["fov"],
                ray_start=kw["ray_start"], ray_end=kw["ray_end"], num_steps=kw["num_steps"], hierarchical_sample=hier,
                clamp_mode=kw["clamp_mode"], noise_std=meta["nerf_noise"], white_back=kw["white_back"], last_back=kw["last_back"],
                lock_view=kw.get("lock_view_dependence", False), debug=True)
    assert rel_err(out["coarse"], ref["coarse"])[0] < 2e-4
    assert rel_err(out["all_z"], ref["all_z"])[0] < 5e-4      # inverse CDF over bins down to 1e-5 wide: conditioning ~1e-4
    assert close_frac(out["rgb"], ref["rgb"], 1e-3)[0] >= 0.995
    assert close_frac(out["depth"][..., None], ref["depth"][..., None], 1e-3)[0] >= 0.995


@pytest.mark.parametrize("n_rays", [25, 7])
def test_emu_pigan_tc_pair_with_a_dummy_partner_group_equals_the_single_cta_kernel(n_rays, monkeypatch):
    """Odd group counts: the last pair runs with a dummy partner group.  Ray-subset renders of the pair kernel must equal the
    full render of the single-CTA kernel bit for bit (same MMA operands, same K order)."""
    sd, z, draws, kw, meta, ref = load_pigan_case("tall_r6_lockview")
    monkeypatch.setenv("C3D_PIGAN_IMPL", "tc")

    def run(n, pair, mode):
        monkeypatch.setenv("C3D_PIGAN_PAIR", "1" if pair else "0")
        with emulated(async_mode=mode, seed=2, sms=2) as pkg:
            G = pkg.pigan.ImplicitGenerator3d(pkg.pigan.TALLSIREN, z_dim=256)
            G.load_state_dict(sd)
            with torch.no_grad():
                fr, ph = pigan_freq_phase(sd, z, meta)
                origin, _, _ = O.camera_origin(draws["yaw_n"], draws["pitch_n"], kw["h_stddev"], kw["v_stddev"], kw["h_mean"], kw["v_mean"])
                return pkg.ops.pigan_render(G.siren.kernel_weights(fr, ph), O.cam2world(-origin, origin), draws["jitter_u"], draws["pdf_u"][:n],
                                            draws["noise_c"][:, :n], draws["noise_f"][:, :n], img_size=meta["img_size"], fov=kw["fov"],
                                            ray_start=kw["ray_start"], ray_end=kw["ray_end"], num_steps=kw["num_steps"],
                                            hierarchical_sample=True, clamp_mode="relu", noise_std=0.0, lock_view=True, n_rays=n, want_depth=True)
    full = run(36, False, 0)
    sub = run(n_rays, True, 2)
    assert torch.equal(sub["rgb"], full["rgb"][:, :n_rays]) and torch.equal(sub["depth"], full["depth"][:, :n_rays])


# ------------------------------------------------------------------ the emulator must report broken kernels (tools/emu/emu_faults.cpp)
@pytest.mark.parametrize("case,mode,needle", [
    (1, "random", "deadlock"), (3, "random", "may only access TMEM lanes"),
    (4, "lazy", "changed between issue and execution"), (5, "lazy", "while a queued tcgen05.mma still writes"),
    (6, "random", "outside the allocation")])
def test_emu_reports_broken_micro_kernels(case, mode, needle, capfd):
    """One deliberately wrong micro-kernel per failure class the emulator claims to detect: a wait nobody satisfies, a
    TMEM lane-quarter violation, an MMA operand overwritten before the MMA ran, a tcgen05.ld of an
    accumulator with the MMA still in flight, an access outside the TMEM allocation.  Case 0 (the same kernel, correct) passes."""
    lib = _emu.emu_lib()
    lib.c3d_emu_fault_case.argtypes = [C.c_int, C.c_void_p]
    out = torch.zeros(128)
    lib.c3d_emu_configure(MODES[mode], 3, 20, 2)
    assert lib.c3d_emu_fault_case(0, out.data_ptr()) == 0 and float(out.abs().sum()) > 0
    assert lib.c3d_emu_fault_case(case, out.data_ptr()) != 0
    assert needle in capfd.readouterr().err


# ------------------------------------------------------------------ CIPS training forward (activation stash) + backward chain
def _cips_chain_fp64(B, N, acts, zs, gp, ws, s1ps, ds, rw, L=18, skip_from=4, rgb_from=3):
    """dZ_l for l = L-1..0 and dL/dx in fp64 from the SAME stash the kernel reads (restates cips_bwd_tc.cu's header)."""
    A = acts.double()
    Wpp = [s1ps[l].double()[:, :, None] * ws[l].double()[None] * ds[l].double()[:, None, :] for l in range(L)]
    dZs, saved, dX = [None] * L, {}, None
    for l in range(L - 1, -1, -1):
        blk, second = l // 2, l % 2 == 1
        G = torch.zeros(B, N, 512, dtype=torch.float64) if l == L - 1 else dX
        if second and blk >= rgb_from:
            G = G + gp.double() @ rw[blk].double()
        if second and (blk + 1) * 2 < L and blk + 1 >= skip_from:
            G = G + saved[blk + 1]
        skip_here = second and blk >= skip_from
        if skip_here:
            saved[blk] = G
            bits = zs[l].to(torch.int32) & 0xFFFF
            m = torch.stack([(bits >> i) & 1 for i in range(16)], -1).reshape(B, N, 512).bool()
        else:
            m = A[l] > 0
        dZs[l] = G * torch.where(m, 1.0, 0.2)
        dX = dZs[l] @ Wpp[l].transpose(1, 2)
    return dZs, dX


@pytest.mark.parametrize("mode", ["eager", "lazy", "random"])
@pytest.mark.parametrize("B,N", [(1, 128), (2, 200)])
def test_emu_cips_backward_chain_matches_fp64_chain_from_the_same_stash(B, N, mode, monkeypatch):
    """c3d_cips_fwd_train (forward + fp16 activation stash + sign bits of z on the residual layers) and c3d_cips_bwd (the
    gradient chain dZ_l = dL/dy_l lrelu'(z_l), dX_l = dZ_l W''_l^T on tcgen05, ToRGB and skip-connection gradients injected in
    the epilogues) against an fp64 evaluation of the same chain from the same stash.  (2, 200): ragged tile, two iterations."""
    monkeypatch.setenv("C3D_CIPS_RES16", "0")       # the training forward keeps the fp32 residual stream; compare like with like
    sd = O.synthetic_state_dict(O.generator_template(), seed=31)
    g = torch.Generator().manual_seed(B * 7 + N)
    x, w, gp = torch.randn(B, N, 32, generator=g), torch.randn(B, 512, generator=g), torch.randn(B, N, 3, generator=g)
    with emulated(async_mode=MODES[mode], seed=N, sms=2) as pkg, torch.no_grad():
        net = build_generator("cpu", sd).inr_net
        ws, s1ps, ds, rw, rb = net.kernel_inputs({k: w for k in net.style_dim_dict}, 9)
        rgb0 = pkg.ops.cips_forward(x, ws, s1ps, ds, rw, rb, impl=TC)
        rgb, acts, zs = pkg.ops.cips_forward_train(x, ws, s1ps, ds, rw, rb)
        dz, dx = pkg.ops.cips_backward_chain(x, acts, zs, gp, ws, s1ps, ds, rw)
    assert torch.equal(rgb, rgb0)                                   # the stash does not change the forward
    dZs, dX = _cips_chain_fp64(B, N, acts, zs, gp, ws, s1ps, ds, rw)
    for l in range(18):
        assert rel_err(dz[l].float(), dZs[l].float())[1] < 2e-3, l     # fp16 operands, 18 layers deep: measured 2e-4 .. 9e-4
    assert rel_err(dx, dX.float())[1] < 2e-3


@pytest.mark.parametrize("pair_env", ["0", "1"])
def test_emu_cips_fused_training_gradients(pair_env, monkeypatch):
    """CIPSNet.train_backend = 'fused' (ops.CipsMLPFunction: native forward + backward chain, weight gradients as fp16 GEMMs
    over the stashes, chain rule into W / modulation / ToRGB in torch) against fp64 autograd of the oracle.  The ToRGB
    gradients (no gate in between) agree to fp16 accuracy; the layer gradients differ by a few per cent in L2 because the
    LeakyReLU gates of the fp16-operand forward differ from the fp64 forward's on the ~0.1 % of units with |z| ~ 0 (each flips
    a gradient factor between 1 and 0.2) -- a property of any reduced-precision forward, not of the backward (the test above
    pins the backward itself at 1e-3).  pair_env = "1": with C3D_CIPS_PAIR set the training forward must still prepare the
    single-CTA kernel's weight tiles (N = 256: two tiles per image, so the inference path WOULD pick the pair kernel) -- a
    mismatch found by dry-running the GPU suite with every variant switched on."""
    monkeypatch.setenv("C3D_CIPS_PAIR", pair_env)
    B, N = 2, 256 if pair_env == "1" else 200
    sd = O.synthetic_state_dict(O.generator_template(), seed=31)
    g = torch.Generator().manual_seed(5)
    x, w, gout = torch.randn(B, N, 32, generator=g), torch.randn(B, 512, generator=g), torch.randn(B, N, 3, generator=g)
    sd64 = {k: (v.double().requires_grad_(k.startswith("inr_net.")) if v.is_floating_point() else v) for k, v in sd.items()}
    x64 = x.double().requires_grad_()
    y_ref = O.cips_net(sd64, x64, w.double())
    (y_ref * gout.double()).sum().backward()
    with emulated(async_mode=2, seed=3, sms=2) as pkg:
        net = build_generator("cpu", sd).train().inr_net
        xr = x.clone().requires_grad_()
        y = net.forward_fused_train(xr, {k: w for k in net.style_dim_dict})
        (y * gout).sum().backward()
    assert rel_err(y.detach(), y_ref.float().detach())[0] < 1e-3
    assert rel_err(xr.grad, x64.grad.float())[1] < 8e-2
    for n, p in net.named_parameters():
        gref = sd64["inr_net." + n].grad
        if gref is None or float(gref.abs().max()) == 0:
            assert p.grad is None or float(p.grad.abs().max()) == 0, n
        else:
            assert rel_err(p.grad, gref.float())[1] < (2e-3 if n.startswith("to_rgbs") else 8e-2), n


def test_emu_fused_style_prep_matches_the_module(monkeypatch):
    """C3D_STYLE_PREP=fused: s1p / demod of all 18 layers from one launch (c3d_cips_style_prep) vs the module's torch ops, and
    the fused forward fed by either."""
    sd = O.synthetic_state_dict(O.generator_template(), seed=31)
    g = torch.Generator().manual_seed(3)
    w, x = torch.randn(3, 512, generator=g), torch.randn(3, 128, 32, generator=g)
    with emulated(async_mode=2, sms=2) as pkg, torch.no_grad():
        net = build_generator("cpu", sd).inr_net
        style = {k: w for k in net.style_dim_dict}
        a = net.kernel_inputs(style, 9)
        monkeypatch.setenv("C3D_STYLE_PREP", "fused")
        b = net.kernel_inputs(style, 9)
        ra = pkg.ops.cips_forward(x, *a, impl=TC)
        rb = pkg.ops.cips_forward(x, *b, impl=TC)
    for l in range(18):
        assert a[1][l].shape == b[1][l].shape and a[2][l].shape == b[2][l].shape
        assert rel_err(b[1][l], a[1][l])[0] < 1e-5 and rel_err(b[2][l], a[2][l])[0] < 1e-5, l
    assert rel_err(rb, ra)[0] < 1e-3        # 1e-7 differences in s1p / demod flip fp16 roundings of the per-image weight tiles


@pytest.mark.parametrize("B,P,Cn,shape3", [(2, 300, 128, True), (1, 5000, 64, False), (3, 17, 256, True), (2, 1, 128, False)])
def test_emu_film_sin_forward_and_backward_match_torch_autograd(B, P, Cn, shape3):
    """csrc/film_ops.cu against the expression it replaces, torch.sin(gain * z + bias) (film_layer.py:94-107), forward and all
    three gradients (dz; dgain / dbias are per-image reductions over the points: partial sums across several chunks)."""
    g = torch.Generator().manual_seed(B * 1000 + P)
    z = torch.randn(B, P, Cn, generator=g).requires_grad_()
    gshape = (B, 1, Cn) if shape3 else (B, Cn)
    gain = (torch.randn(*gshape, generator=g) * 5 + 30).requires_grad_()
    bias = torch.randn(*gshape, generator=g).requires_grad_()
    dy = torch.randn(B, P, Cn, generator=g)
    ref = torch.sin(gain.view(B, 1, Cn).double() * z.double() + bias.view(B, 1, Cn).double())
    rz, rg, rb = torch.autograd.grad(ref, (z, gain, bias), dy.double())
    with emulated(async_mode=0) as pkg:
        assert pkg.ops.film_sin_supported(z, gain, bias)
        y = pkg.ops.film_sin(z, gain, bias)
        dz, dg, db = torch.autograd.grad(y, (z, gain, bias), dy)
    assert dg.shape == gain.shape and db.shape == bias.shape
    assert (y.double() - ref).abs().max().item() < 2e-5               # fp32 argument of magnitude ~30-100: |d arg| ~ 4e-6
    assert (dz.double() - rz).abs().max().item() < 2e-5 * rz.abs().max().item() + 1e-6
    assert (dg.double() - rg).abs().max().item() < 1e-4 * rg.abs().max().item() + 1e-5
    assert (db.double() - rb).abs().max().item() < 1e-4 * rb.abs().max().item() + 1e-5


def test_emu_film_layer_fused_flag_matches_torch_path():
    """FiLMLayer.fused_film: same module, same parameters; outputs and parameter gradients of the native op vs the torch ops."""
    import cips3d_b200
    torch.manual_seed(3)
    layer = cips3d_b200.FiLMLayer(3, 128, 64)
    x, style = torch.randn(2, 257, 3), torch.randn(2, 64)
    out = {}
    for fused in (False, True):
        layer.fused_film = fused
        layer.zero_grad()
        with emulated(async_mode=0):
            y = layer(x, style)
            y.square().sum().backward()
        out[fused] = (y.detach().clone(), {k: v.grad.clone() for k, v in layer.named_parameters()})
    assert (out[True][0] - out[False][0]).abs().max().item() < 2e-5
    for k, gr in out[False][1].items():
        assert (out[True][1][k] - gr).abs().max().item() < 2e-4 * gr.abs().max().item() + 1e-5, k
    with emulated(async_mode=0) as pkg:     # unsupported widths fall back to the torch expression
        assert not pkg.ops.film_sin_supported(torch.zeros(1, 4, 6), torch.zeros(1, 6), torch.zeros(1, 6))


def _pigan_render(pkg, name):
    sd, z, draws, kw, meta, ref = load_pigan_case(name)
    cls = pkg.pigan.SPATIALSIRENBASELINE if meta["siren_cls"] == "SPATIALSIRENBASELINE" else pkg.pigan.TALLSIREN
    G = pkg.pigan.ImplicitGenerator3d(cls, z_dim=256)
    G.load_state_dict(sd)
    hier = kw["hierarchical_sample"]
    with torch.no_grad():
        fr, ph = pigan_freq_phase(sd, z, meta)
        origin, _, _ = O.camera_origin(draws["yaw_n"], draws["pitch_n"], kw["h_stddev"], kw["v_stddev"], kw["h_mean"], kw["v_mean"])
        out = pkg.ops.pigan_render(
            G.siren.kernel_weights(fr, ph), O.cam2world(-origin, origin), draws["jitter_u"], draws["pdf_u"] if hier else None,
            draws["noise_c"] if hier else None, draws["noise_f"], img_size=meta["img_size"], fov=kw["fov"],
            ray_start=kw["ray_start"], ray_end=kw["ray_end"], num_steps=kw["num_steps"], hierarchical_sample=hier,
            clamp_mode=kw["clamp_mode"], noise_std=meta["nerf_noise"], white_back=kw["white_back"], last_back=kw["last_back"],
            lock_view=kw.get("lock_view_dependence", False), debug=True)
    return out, ref


_SCHEDULES = [dict(async_mode=0, seed=1, preempt_permille=0), dict(async_mode=1, seed=2, preempt_permille=100),
              dict(async_mode=2, seed=99, preempt_permille=300, sms=3)]


def _bits(tensors):
    import hashlib
    return tuple(hashlib.sha1(t.detach().contiguous().numpy().tobytes()).hexdigest() for t in tensors)


@pytest.mark.parametrize("what,env", [("ray", {}), ("ray", {"C3D_RAY_MATH": "fold"}), ("cips", {"C3D_CIPS_PAIR": "1"}), ("cips_image", {"C3D_CIPS_PAIR": "1"}),
                                      ("cips_train", {}), ("pigan", {"C3D_PIGAN_IMPL": "tc", "C3D_PIGAN_PAIR": "1"})])
def test_emu_outputs_are_bit_identical_across_schedules(what, env, monkeypatch):
    """A race that stays inside the parity tolerances would still make the output depend on WHEN asynchronous operations
    complete, which thread runs first and which CTA gets which tile.  Three very different schedules (eager / latest-possible /
    random completion, 0-30 % random preemption, 2 or 3 SMs) must give the same bits -- for every kernel family, default and
    opt-in forms.  (It also shows the CTA-pair CIPS kernel equal to the single-CTA kernel bit for bit: see the assertion.)"""
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    sd = O.synthetic_state_dict(O.generator_template(), seed=31)
    g = torch.Generator().manual_seed(3)
    x, w = torch.randn(2, 512, 32, generator=g), torch.randn(2, 512, generator=g)

    def run(pkg):
        if what == "ray":
            out, _ = _render(pkg, "r16_trained_noise", TC, debug=False, want_depth=True, want_weights=True)
            return out["pixels_fea"], out["depth"], out["weights"]
        if what in ("cips", "cips_image"):     # cips_image: no hidden state asked for -> fp16 residual stream
            G = build_generator("cpu", sd)
            with torch.no_grad():
                ws, s1p, dm, rw, rb = G.inr_net.kernel_inputs({k: w for k in G.inr_net.style_dim_dict}, 9)
                if what == "cips_image":
                    return (pkg.ops.cips_forward(x, ws, s1p, dm, rw, rb, impl=TC),)
                return pkg.ops.cips_forward(x, ws, s1p, dm, rw, rb, impl=TC, return_hidden=True)
        if what == "cips_train":
            net = build_generator("cpu", sd).train().inr_net
            xr = x.clone().requires_grad_()
            y = net.forward_fused_train(xr, {k: w for k in net.style_dim_dict})
            y.square().sum().backward()
            return (y, xr.grad) + tuple(p.grad for p in net.parameters() if p.grad is not None)
        out, _ = _pigan_render(pkg, "spatial_r8_noise_backs")
        return tuple(out[k] for k in sorted(out) if torch.is_tensor(out[k]))

    seen = []
    for s in _SCHEDULES:
        with emulated(**s) as pkg:
            seen.append(_bits(run(pkg)))
    assert seen[0] == seen[1] == seen[2]
    if what in ("cips", "cips_image"):          # and the pair form computes exactly what the single-CTA form computes
        monkeypatch.setenv("C3D_CIPS_PAIR", "0")
        with emulated(**_SCHEDULES[0]) as pkg:
            assert _bits(run(pkg)) == seen[0]


# ------------------------------------------------------------------ volume integration of the autograd graph (csrc/integrate_ops.cu)
from _integrate_cases import (CASES as INTEG_CASES, GOLDEN_CASES as INTEG_GOLDEN, MERGED_CASES, PDF_GOLDEN_CASES,  # noqa: E402
                              check as integ_check, check_golden, check_merged, check_pdf_golden, make as integ_make,
                              make_merged)


@pytest.mark.parametrize("name", PDF_GOLDEN_CASES)
def test_emu_sample_pdf_matches_golden_of_the_real_function(name):
    """c3d_sample_pdf against outputs of the UNMODIFIED pigan_utils.sample_pdf (12 / 32 / 22 bins, 12 / 40 / 24 draws, det mode,
    empty bins that hit the denom < eps branch)."""
    with emulated(async_mode=0) as pkg:
        check_pdf_golden(name, pkg)
        with pytest.raises(pkg._lib.C3dError, match="unsupported shapes"):
            pkg.ops.sample_pdf_from_u(torch.zeros(2, 35), torch.zeros(2, 34), torch.zeros(2, 4))


@pytest.mark.parametrize("name", INTEG_GOLDEN)
def test_emu_integrate_matches_golden_of_the_real_fancy_integration(name):
    """tests/golden/fancy_integration.npz holds what the UNMODIFIED pigan_utils.fancy_integration returned (and its autograd
    gradient) -- for the merged cases after the reference's own cat + sort + gather; the native op must reproduce it."""
    with emulated(async_mode=0) as pkg:
        check_golden(name, pkg)


@pytest.mark.parametrize("idx", range(len(INTEG_CASES)))
def test_emu_integrate_forward_and_backward_match_fp64_autograd(idx):
    """ops.IntegrateFunction against fp64 torch autograd of the oracle's fancy_integration (pigan_utils.py:212-273): features,
    weights and the gradient w.r.t. every (feature_i, sigma_i), for both clamps, last_back / white_back, noise, 1..32 samples."""
    case = INTEG_CASES[idx]
    rs, z, noise, d_fea = integ_make(case, seed=100 + idx)
    with emulated(async_mode=0) as pkg:
        integ_check(case, pkg, rs, z, noise, d_fea)


@pytest.mark.parametrize("idx", range(len(MERGED_CASES)))
def test_emu_integrate_merged_matches_fp64_autograd_of_cat_sort_gather_integrate(idx):
    """ops.IntegrateMergedFunction: unsorted fine + coarse halves in, generator.py:1733-1752 (cat, sort, gather, fancy_integration)
    out; the sorted depths bit for bit (incl. a tie across the halves), weights, features, and the gradients at the SOURCE rows."""
    case = MERGED_CASES[idx]
    with emulated(async_mode=0) as pkg:
        check_merged(case, pkg, *make_merged(case, seed=200 + idx))


def test_emu_integrate_matches_fp32_torch_ops_it_replaces():
    """Same dtype, same formulas: the op vs generator._torch_integrate (what the training graph ran before) incl. relu'(0) = 0."""
    import cips3d_b200
    case = INTEG_CASES[0]
    rs, z, noise, d_fea = integ_make(case, seed=7)
    rs[0, 0, 3, 32] = 0.0                                                   # sigma exactly 0: relu' = 0 on both sides
    rs[0, 1, -1, 32] = -1.0                                                 # last sample empty: delta = 1e10 * relu(-1) = 0
    r0 = rs.clone().requires_grad_()
    fea0, w0 = cips3d_b200.generator._torch_integrate(r0, z, None, "relu", True, False, 32)
    (d0,) = torch.autograd.grad(fea0, r0, d_fea)
    with emulated(async_mode=0) as pkg:
        r1 = rs.clone().requires_grad_()
        fea1, w1 = pkg.ops.integrate(r1, z, None, "relu", True, False)
        (d1,) = torch.autograd.grad(fea1, r1, d_fea)
    assert (w1 - w0).abs().max().item() < 1e-6 and (fea1 - fea0).abs().max().item() < 1e-5
    assert d1[0, 0, 3, 32].item() == 0.0 and d1[0, 1, -1, 32].item() == 0.0
    assert (d1 - d0).abs().max().item() < 1e-4 * d0.abs().max().item() + 1e-6


def test_emu_integrate_rejects_unsupported_shapes_and_generator_flag(monkeypatch):
    """More than 32 samples per ray: integrate_supported says no and the generator keeps the torch ops; the flag on the real
    generator's training graph gives the same image and the same parameter gradients as the torch ops."""
    import cips3d_b200
    from oracle import cips3d_oracle as O2
    monkeypatch.setattr(cips3d_b200.generator, "_require_cuda", lambda *a, **k: None)
    with emulated(async_mode=0) as pkg:
        assert not pkg.ops.integrate_supported(torch.zeros(2, 3, 33, 33), torch.zeros(2, 3, 33))
        assert not pkg.ops.integrate_supported(torch.zeros(2, 3, 8, 33), torch.zeros(2, 3, 9))
        with pytest.raises(pkg._lib.C3dError, match="samples per ray"):
            pkg.ops.integrate(torch.zeros(1, 2, 40, 33), torch.zeros(1, 2, 40))
        G = build_generator("cpu", O2.synthetic_state_dict(O2.generator_template(), seed=5, sigma_bias=0.3)).train()
        kw = dict(O2.G_KWARGS)
        kw["num_steps"] = 6
        res = {}
        for backend in ("torch", "fused"):
            G.train_integrate = backend
            G.zero_grad()
            torch.manual_seed(11)
            img, _ = G(G.get_zs(2), img_size=8, nerf_noise=0.5, **kw)
            img.square().mean().backward()
            res[backend] = (img.detach().clone(), {k: p.grad.clone() for k, p in G.named_parameters() if p.grad is not None})
    assert (res["fused"][0] - res["torch"][0]).abs().max().item() < 1e-4
    assert res["fused"][1].keys() == res["torch"][1].keys() and any(k.startswith("siren.") for k in res["fused"][1])
    for k, gr in res["torch"][1].items():
        assert (res["fused"][1][k] - gr).abs().max().item() < 1e-3 * gr.abs().max().item() + 1e-7, k


@pytest.mark.parametrize("name", ["spatial_r8", "spatial_r8_noise_backs", "tall_r6_lockview", "spatial_r6_nohier_s24"])
def test_emu_pigan_training_graph_with_native_integration_matches_reference_golden(name):
    """ImplicitGenerator3d.train_integrate = 'fused' (rgb_dim 3: one partial channel group; coarse weights under no_grad and the
    final compositing): the image of the REAL piGAN_lib classes' golden, and the field gradients of the torch-op graph."""
    import cips3d_b200
    sd, z, draws, kw, meta, ref = load_pigan_case(name)
    cls = cips3d_b200.pigan.SPATIALSIRENBASELINE if meta["siren_cls"] == "SPATIALSIRENBASELINE" else cips3d_b200.pigan.TALLSIREN
    G = cips3d_b200.pigan.ImplicitGenerator3d(cls, z_dim=256)
    G.load_state_dict(sd)
    hier = kw["hierarchical_sample"]
    origin, _, _ = O.camera_origin(draws["yaw_n"], draws["pitch_n"], kw["h_stddev"], kw["v_stddev"], kw["h_mean"], kw["v_mean"])
    res = {}
    for backend in ("torch", "fused"):
        G.train_integrate = backend
        G.zero_grad()
        with emulated(async_mode=0):
            fr, ph = G.siren.mapping_network(z)
            rgb, depth = G._render_torch(fr, ph, O.cam2world(-origin, origin), draws["jitter_u"], draws["pdf_u"] if hier else None,
                                         draws["noise_c"] if hier else None, draws["noise_f"], meta["img_size"], kw["fov"],
                                         kw["ray_start"], kw["ray_end"], kw["num_steps"], hier, kw.get("lock_view_dependence", False),
                                         kw["clamp_mode"], meta["nerf_noise"], kw["white_back"], kw["last_back"])
            img = G._to_img(rgb, meta["img_size"])
            img.square().mean().backward()
        res[backend] = (img.detach().clone(), depth.detach().clone(), {k: p.grad.clone() for k, p in G.named_parameters()})
    assert rel_err(res["fused"][0], ref["img"])[0] < 1e-4 and rel_err(res["fused"][1], ref["depth"])[0] < 1e-4
    assert (res["fused"][0] - res["torch"][0]).abs().max().item() < 2e-5
    for k, gr in res["torch"][2].items():
        assert (res["fused"][2][k] - gr).abs().max().item() < 1e-3 * gr.abs().max().item() + 1e-8, k


@pytest.mark.parametrize("backend", ["torch", "fused"])
@pytest.mark.parametrize("name", ["hier_noise_lastback", "flat_softplus_white"])
def test_emu_points_forward_matches_golden_of_the_real_methods(name, backend, monkeypatch):
    """comm_utils.get_world_points_and_direction -> GeneratorNerfINR.points_forward with the reference's recorded draws replayed:
    the seven world tensors and the inr / aux images the UNMODIFIED reference produced (tests/golden/points_forward.npz)."""
    import cips3d_b200
    from _integrate_cases import check_points_golden
    monkeypatch.setattr(cips3d_b200.generator, "_require_cuda", lambda *a, **k: None)
    G = build_generator("cpu", O.synthetic_state_dict(O.generator_template(), seed=77, sigma_bias=0.3))
    with emulated(async_mode=0) as pkg:
        err = check_points_golden(name, pkg, G, "cpu", backend)
    assert err < 2e-4             # the CIPS MLP runs on the fused fp16 tensor-core kernel here (no graph needed)


@pytest.mark.parametrize("sample_dist", [None, "mean", "not_a_mode"])
def test_emu_pigan_unknown_sample_dist_is_the_mean_camera(sample_dist, monkeypatch):
    """piGAN_lib/generators/volumetric_rendering.py:162-165: any mode sample_camera_positions does not know -- None included,
    the default of ImplicitGenerator3d.forward / staged_forward and what inverse_render.py passes -- renders from
    (h_mean, v_mean).  (The CIPS-3D surface asserts on unknown modes)"""
    import cips3d_b200
    monkeypatch.setattr(cips3d_b200.pigan, "_require_cuda", lambda *a, **k: None)
    torch.manual_seed(3)
    G = cips3d_b200.pigan.ImplicitGenerator3d(cips3d_b200.pigan.TALLSIREN, z_dim=256).eval()
    z = torch.randn(2, 256)
    kw = dict(img_size=4, fov=12, ray_start=0.88, ray_end=1.12, num_steps=4, h_stddev=0.3, v_stddev=0.155, h_mean=1.3, v_mean=1.7,
              hierarchical_sample=False, clamp_mode="relu", nerf_noise=0.0)
    with torch.no_grad(), emulated(async_mode=0):
        img, py = G(z, sample_dist=sample_dist, **kw)
    assert torch.isfinite(img).all()
    assert torch.allclose(py, torch.tensor([[1.7, 1.3]]).expand(2, 2), atol=1e-6)       # (pitch, yaw) = (v_mean, h_mean)


@pytest.mark.parametrize("shape,pad", [((2, 8, 16, 16), (2, 2, 2, 2)), ((1, 128, 9, 7), (1, 1, 1, 1)), ((2, 40, 33, 5), (2, 1, 0, 3)),
                                       ((1, 3, 4, 4), (3, 3, 3, 3)), ((1, 16, 37, 65), (-1, 2, 2, -1))])
def test_emu_channels_last_discriminator_ops_match_nchw(shape, pad):
    """ops.bias_act and the 4x4 blur on channels-last tensors (c3d_blur_nhwc: the register-streaming kernel with the taps C
    elements apart) equal the NCHW results element for element, keep the layout, and differentiate twice the same way."""
    g = torch.Generator().manual_seed(sum(shape))
    x = torch.randn(*shape, generator=g)
    b = torch.randn(shape[1], generator=g)
    k1 = torch.tensor([1., 3., 3., 1.])
    xcl = x.contiguous(memory_format=torch.channels_last)
    for k in ((k1[None] * k1[:, None]) / 64, torch.randn(4, 4, generator=g)):
        with emulated(async_mode=0, sms=2) as pkg:
            y0 = pkg.ops._upfirdn2d_raw(x, k, (1, 1), (1, 1), pad)
            y1 = pkg.ops._upfirdn2d_raw(xcl, k, (1, 1), (1, 1), pad)
            a0 = pkg.ops.bias_act(x, b)
            a1 = pkg.ops.bias_act(xcl, b)
            g0 = pkg.ops.bias_act(x, None, a0, 3, 1)
            g1 = pkg.ops.bias_act(xcl, None, a0, 3, 1)          # NCHW ref against a channels-last input: converted inside
        assert y1.shape == y0.shape and (shape[1] == 1 or y1.is_contiguous(memory_format=torch.channels_last))
        assert torch.equal(a1, a0) and torch.equal(g1, g0) and a1.stride() == xcl.stride()
        assert (y1 - y0).abs().max().item() < 1e-6
        assert (y0 - O.upfirdn2d(x, k, (1, 1), (1, 1), pad)).abs().max().item() < 1e-5
    with emulated(async_mode=0, sms=2) as pkg:          # autograd through both ops, R1 pattern, both layouts
        kk = (k1[None] * k1[:, None]) / 64
        res = []
        for inp in (x, xcl):
            xi = inp.clone().requires_grad_()
            bi = b.clone().requires_grad_()
            y = pkg.ops.fused_leaky_relu(pkg.ops.upfirdn2d(xi, kk, pad=(max(pad[0], 0), max(pad[1], 0))), bi)
            gx, = torch.autograd.grad(y.square().sum(), xi, create_graph=True)
            gx.square().sum().backward()
            res.append((y.detach(), gx.detach(), xi.grad, bi.grad))
        for a, c in zip(*res):
            assert (a - c).abs().max().item() <= 1e-5 * max(1.0, c.abs().max().item())


@pytest.mark.parametrize("rows,K,N,bias", [(128, 128, 128, True), (300, 128, 64, True), (77, 64, 32, True), (1, 32, 64, False),
                                            (513, 64, 128, False), (256, 128, 32, True)])
@pytest.mark.parametrize("mode", ["eager", "random"])
def test_emu_points_linear_matches_fp64(rows, K, N, bias, mode):
    """c3d_points_linear (tcgen05 split-fp16 GEMM: the NeRF field's per-point linears in the training graph) against an fp64
    product: forward, the transposed form (data gradient) with the device-side gradient scale, ragged last tile, several tiles
    per CTA (2 emulated SMs), and the autograd Function against torch's linear."""
    g = torch.Generator().manual_seed(rows * 7 + K + N)
    x = torch.randn(rows, K, generator=g)
    w = torch.randn(N, K, generator=g) * 0.3
    b = torch.randn(N, generator=g) if bias else None
    with emulated(async_mode=MODES[mode], seed=3, sms=2) as pkg:
        y = pkg.ops._points_linear_raw(x, w, b, None, False)
        dz = torch.randn(rows, N, generator=g) * 1e-6                      # gradient-sized operand: needs the scale
        sc = (1024.0 / dz.abs().amax()).reshape(1)
        dx = pkg.ops._points_linear_raw(dz, w, None, sc, True)
        xr = x.view(1, rows, K).clone().requires_grad_()
        wr, br = w.clone().requires_grad_(), (b.clone().requires_grad_() if bias else None)
        out = pkg.ops.points_linear(xr, wr, br)
        out.square().sum().backward()
    ref = x.double() @ w.double().T + (b.double() if bias else 0)
    assert (y.double() - ref).abs().max().item() < 2e-6 * ref.abs().max().item()
    refx = dz.double() @ w.double()
    assert (dx.double() - refx).abs().max().item() < 2e-6 * refx.abs().max().item()
    x2 = x.view(1, rows, K).clone().requires_grad_()
    w2, b2 = w.clone().requires_grad_(), (b.clone().requires_grad_() if bias else None)
    torch.nn.functional.linear(x2, w2, b2).square().sum().backward()
    assert (xr.grad - x2.grad).abs().max().item() < 1e-5 * x2.grad.abs().max().item()
    assert (wr.grad - w2.grad).abs().max().item() < 1e-4 * w2.grad.abs().max().item()
    if bias:
        assert (br.grad - b2.grad).abs().max().item() < 1e-4 * b2.grad.abs().max().item()
