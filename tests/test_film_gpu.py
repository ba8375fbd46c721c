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
