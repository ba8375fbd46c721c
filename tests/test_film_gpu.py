"""FiLM + sine autograd op (csrc/film_ops.cu) on the GPU against the torch expression it replaces, same device."""
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.fixture(scope="module")
def pkg():
    import cips3d_b200
    cips3d_b200._lib.load()
    return cips3d_b200


@pytest.mark.parametrize("B,P,Cn", [(4, 65536 * 6, 128), (2, 4097, 64), (1, 1000, 256), (3, 1, 128)])
def test_film_sin_matches_torch(pkg, B, P, Cn):
    g = torch.Generator().manual_seed(P)
    z = torch.randn(B, P, Cn, generator=g).to(DEV).requires_grad_()
    gain = (torch.randn(B, 1, Cn, generator=g) * 5 + 30).to(DEV).requires_grad_()
    bias = torch.randn(B, 1, Cn, generator=g).to(DEV).requires_grad_()
    dy = torch.randn(B, P, Cn, generator=g).to(DEV)
    ref = torch.sin(gain * z + bias)                                   # film_layer.py:107
    rz, rg, rb = torch.autograd.grad(ref, (z, gain, bias), dy)
    y = pkg.ops.film_sin(z, gain, bias)
    dz, dg, db = torch.autograd.grad(y, (z, gain, bias), dy)
    assert (y - ref).abs().max().item() < 1e-6                        # same mul / add / sinf sequence as torch's kernels
    assert (dz - rz).abs().max().item() < 1e-5 * rz.abs().max().item() + 1e-6
    # reductions over up to 393 216 points in fp32: both sides carry ~1e-6 relative summation error
    assert (dg - rg).abs().max().item() < 2e-4 * rg.abs().max().item() + 1e-4
    assert (db - rb).abs().max().item() < 2e-4 * rb.abs().max().item() + 1e-4


def test_film_layer_flag_in_the_generator_training_graph(pkg):
    """GeneratorNerfINR with fused_film on every NeRF FiLMLayer: same image and same parameter gradients as the torch ops."""
    from _util import build_generator
    from oracle import cips3d_oracle as O
    G = build_generator(DEV, O.synthetic_state_dict(O.generator_template(), seed=5, sigma_bias=0.3)).train()
    kw = dict(O.G_KWARGS)
    kw["num_steps"] = 6
    res = {}
    for fused in (False, True):
        for m in G.modules():
            if isinstance(m, pkg.FiLMLayer):
                m.fused_film = fused
        G.zero_grad()
        torch.manual_seed(11)
        img, _ = G(G.get_zs(2), img_size=16, nerf_noise=0.0, **kw)
        img.square().mean().backward()
        res[fused] = (img.detach().clone(), {k: p.grad.clone() for k, p in G.named_parameters() if p.grad is not None})
    assert (res[True][0] - res[False][0]).abs().max().item() < 1e-4
    assert res[True][1].keys() == res[False][1].keys() and any(k.startswith("siren.") for k in res[True][1])
    for k, gr in res[False][1].items():
        assert (res[True][1][k] - gr).abs().max().item() < 1e-3 * gr.abs().max().item() + 1e-7, k


@pytest.mark.parametrize("rows,K,N,bias", [(65536 * 24, 128, 128, True), (100003, 128, 64, True), (4097, 64, 32, True), (1, 32, 64, False),
                                            (70000, 64, 128, False), (513, 32, 32, True)])
def test_points_linear_matches_fp64(pkg, rows, K, N, bias):
    """c3d_points_linear (tcgen05 split-fp16 GEMM of the NeRF training graph) against an fp64 product: forward, the transposed
    (data gradient) form with gradient-sized operands and the device-side scale, and the autograd Function vs torch's linear."""
    g = torch.Generator().manual_seed(rows % 1000 + K + N)
    x = torch.randn(rows, K, generator=g).to(DEV)
    w = (torch.randn(N, K, generator=g) * 0.3).to(DEV)
    b = torch.randn(N, generator=g).to(DEV) if bias else None
    y = pkg.ops._points_linear_raw(x, w, b, None, False)
    n_chk = min(rows, 20000)
    ref = x[:n_chk].double() @ w.double().T + (b.double() if bias else 0)
    assert (y[:n_chk].double() - ref).abs().max().item() < 2e-6 * ref.abs().max().item()
    dz = torch.randn(rows, N, generator=g).to(DEV) * 1e-6
    sc = (1024.0 / dz.abs().amax()).reshape(1)
    dx = pkg.ops._points_linear_raw(dz, w, None, sc, True)
    refx = dz[:n_chk].double() @ w.double()
    assert (dx[:n_chk].double() - refx).abs().max().item() < 2e-6 * refx.abs().max().item()
    if rows <= 100003:
        xr = x.view(1, rows, K).clone().requires_grad_()
        wr, br = w.clone().requires_grad_(), (b.clone().requires_grad_() if bias else None)
        pkg.ops.points_linear(xr, wr, br).square().sum().backward()
        x2 = x.view(1, rows, K).clone().requires_grad_()
        w2, b2 = w.clone().requires_grad_(), (b.clone().requires_grad_() if bias else None)
        torch.nn.functional.linear(x2, w2, b2).square().sum().backward()
        assert (xr.grad - x2.grad).abs().max().item() < 1e-5 * x2.grad.abs().max().item()
        assert (wr.grad - w2.grad).abs().max().item() < 1e-4 * w2.grad.abs().max().item()


def test_points_linear_in_the_generator_training_graph(pkg):
    """GeneratorNerfINR with every hot op of the NeRF training graph native (FiLM + sine, per-point linears, volume integration):
    same image and parameter gradients as the torch-op graph."""
    from _util import build_generator
    from oracle import cips3d_oracle as O
    G = build_generator(DEV, O.synthetic_state_dict(O.generator_template(), seed=5, sigma_bias=0.3)).train()
    kw = dict(O.G_KWARGS)
    kw["num_steps"] = 6
    res = {}
    for fused in (False, True):
        for m in G.modules():
            if isinstance(m, pkg.FiLMLayer):
                m.fused_film = m.fused_linear = fused
        G.siren.fused_linear = fused
        G.train_integrate = "fused" if fused else "torch"
        G.zero_grad()
        torch.manual_seed(11)
        img, _ = G(G.get_zs(2), img_size=16, nerf_noise=0.0, **kw)
        img.square().mean().backward()
        res[fused] = (img.detach().clone(), {k: p.grad.clone() for k, p in G.named_parameters() if p.grad is not None})
    assert (res[True][0] - res[False][0]).abs().max().item() < 1e-4
    assert res[True][1].keys() == res[False][1].keys() and any(k.startswith("siren.") for k in res[True][1])
    for k, gr in res[False][1].items():
        assert (res[True][1][k] - gr).abs().max().item() < 1e-3 * gr.abs().max().item() + 1e-7, k
