"""Volume integration op (csrc/integrate_ops.cu) on the GPU: against fp64 autograd of the oracle (the cases of the emulation
tests), against the torch CUDA ops it replaces at a training-sized batch, and inside the generator's training graph."""
import pytest
import torch

from _integrate_cases import (CASES, GOLDEN_CASES, MERGED_CASES, PDF_GOLDEN_CASES, check, check_golden, check_merged,
                              check_pdf_golden, make, make_merged)

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.fixture(scope="module")
def pkg():
    import cips3d_b200
    cips3d_b200._lib.load()
    return cips3d_b200


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_integrate_matches_golden_of_the_real_fancy_integration(pkg, name):
    check_golden(name, pkg, DEV)


@pytest.mark.parametrize("name", PDF_GOLDEN_CASES)
def test_sample_pdf_matches_golden_of_the_real_function(pkg, name):
    check_pdf_golden(name, pkg, DEV)


@pytest.mark.parametrize("idx", range(len(CASES)))
def test_integrate_matches_fp64_autograd(pkg, idx):
    rs, z, noise, d_fea = make(CASES[idx], seed=100 + idx, device=DEV)
    check(CASES[idx], pkg, rs, z, noise, d_fea)


@pytest.mark.parametrize("idx", range(len(MERGED_CASES)))
def test_integrate_merged_matches_fp64_autograd(pkg, idx):
    check_merged(MERGED_CASES[idx], pkg, *make_merged(MERGED_CASES[idx], seed=200 + idx, device=DEV))


def test_integrate_merged_training_size_matches_the_torch_cuda_graph(pkg):
    """2 x 64 x 64 rays, 12 + 12 samples: torch's cat + sort + gather + integrate on the same GPU, forward and both gradients."""
    case = ((2, 4096), 12, 32, "relu", True, False, False)
    fine, zf, coarse, zc, _, d_fea = make_merged(case, seed=4, device=DEV)
    f0, c0 = fine.clone().requires_grad_(), coarse.clone().requires_grad_()
    all_z, ind = torch.sort(torch.cat([zf, zc], -1), dim=-1, stable=True)
    all_out = torch.gather(torch.cat([f0, c0], -2), -2, ind[..., None].expand(-1, -1, -1, 33))
    fea0, w0 = pkg.generator._torch_integrate(all_out, all_z, None, "relu", True, False, 32)
    g0 = torch.autograd.grad(fea0, (f0, c0), d_fea)
    f1, c1 = fine.clone().requires_grad_(), coarse.clone().requires_grad_()
    fea1, w1, z1 = pkg.ops.integrate_merged(f1, zf, c1, zc, None, "relu", True, False)
    g1 = torch.autograd.grad(fea1, (f1, c1), d_fea)
    assert torch.equal(z1, all_z)
    assert (w1 - w0).abs().max().item() < 2e-6 and (fea1 - fea0).abs().max().item() < 2e-5
    for a, b in zip(g1, g0):
        assert (a - b).abs().max().item() < 2e-4 * max(g.abs().max().item() for g in g0) + 1e-6


def test_integrate_training_size_matches_torch_cuda_ops(pkg):
    """2 x 64 x 64 rays, 24 sorted samples (more rays than resident warps: the grid-stride loop runs), vs _torch_integrate."""
    case = ((2, 4096), 24, 32, "relu", True, False, False)
    rs, z, _, d_fea = make(case, seed=3, device=DEV)
    r0 = rs.clone().requires_grad_()
    fea0, w0 = pkg.generator._torch_integrate(r0, z, None, "relu", True, False, 32)
    (d0,) = torch.autograd.grad(fea0, r0, d_fea)
    r1 = rs.clone().requires_grad_()
    fea1, w1 = pkg.ops.integrate(r1, z, None, "relu", True, False)
    (d1,) = torch.autograd.grad(fea1, r1, d_fea)
    assert (w1 - w0).abs().max().item() < 2e-6 and (fea1 - fea0).abs().max().item() < 2e-5
    assert (d1 - d0).abs().max().item() < 2e-4 * d0.abs().max().item() + 1e-6
    assert (w1.sum(-1) - 1).abs().max().item() < 1e-5                      # last_back: the weights of a ray sum to one


def test_generator_flag_in_the_training_graph(pkg):
    """GeneratorNerfINR.train_integrate = 'fused': same image and parameter gradients as the torch ops."""
    from _util import build_generator
    from oracle import cips3d_oracle as O
    G = build_generator(DEV, O.synthetic_state_dict(O.generator_template(), seed=5, sigma_bias=0.3)).train()
    kw = dict(O.G_KWARGS)
    kw["num_steps"] = 6
    res = {}
    for backend in ("torch", "fused"):
        G.train_integrate = backend
        G.zero_grad()
        torch.manual_seed(11)
        img, _ = G(G.get_zs(2), img_size=16, nerf_noise=0.5, **kw)
        img.square().mean().backward()
        res[backend] = (img.detach().clone(), {k: p.grad.clone() for k, p in G.named_parameters() if p.grad is not None})
    assert (res["fused"][0] - res["torch"][0]).abs().max().item() < 1e-4
    assert res["fused"][1].keys() == res["torch"][1].keys() and any(k.startswith("siren.") for k in res["fused"][1])
    for k, gr in res["torch"][1].items():
        assert (res["fused"][1][k] - gr).abs().max().item() < 1e-3 * gr.abs().max().item() + 1e-7, k


@pytest.mark.parametrize("frozen", [False, True])
def test_points_forward_native_ops_match_torch_ops(pkg, frozen):
    """get_world_points_and_direction -> points_forward (reference entry points, generator.py:1659-1762) on the GPU: the native
    integration / resampling / merge ops against the torch ops on the same draws -- images, aux images, parameter gradients."""
    from _util import build_generator
    from oracle import cips3d_oracle as O
    G = build_generator(DEV, O.synthetic_state_dict(O.generator_template(), seed=5, sigma_bias=0.3), frozen=frozen).train()
    torch.manual_seed(3)
    zs = G.get_zs(2)
    res = {}
    for backend in ("torch", "fused"):
        G.train_integrate = backend
        G.zero_grad()
        torch.manual_seed(17)
        pts, dirs_exp, origins, dirs, z_vals, pitch, yaw = pkg.comm_utils.get_world_points_and_direction(
            batch_size=2, num_steps=12, img_size=16, fov=12, ray_start=0.88, ray_end=1.12, h_stddev=0.3, v_stddev=0.155,
            h_mean=1.5707963, v_mean=1.5707963, sample_dist="gaussian", lock_view_dependence=False, device=DEV)
        style = G.mapping_network(**zs)
        inr, aux = G.points_forward(style_dict=style, transformed_points=pts.view(2, 256, 12, 3),
                                    transformed_ray_directions_expanded=dirs_exp.view(2, 256, 12, 3), num_steps=12,
                                    hierarchical_sample=True, z_vals=z_vals, clamp_mode="relu", nerf_noise=0.3,
                                    transformed_ray_origins=origins, transformed_ray_directions=dirs, white_back=False,
                                    last_back=True, return_aux_img=True, idx_grad=torch.arange(0, 256, 2, device=DEV))
        assert inr.shape == (2, 128, 3) and aux.shape == (2, 128, 3)
        (inr.square().mean() + (aux.square().mean() if aux.requires_grad else 0)).backward()
        res[backend] = (inr.detach().clone(), aux.detach().clone(),
                        {k: p.grad.clone() for k, p in G.named_parameters() if p.grad is not None})
    assert (res["fused"][0] - res["torch"][0]).abs().max().item() < 1e-4 and (res["fused"][1] - res["torch"][1]).abs().max().item() < 1e-4
    assert res["fused"][2].keys() == res["torch"][2].keys()
    assert any(k.startswith("siren.") for k in res["torch"][2]) == (not frozen)
    for k, gr in res["torch"][2].items():
        assert (res["fused"][2][k] - gr).norm().item() < 1e-2 * gr.norm().item() + 1e-7, k


@pytest.mark.parametrize("backend", ["torch", "fused"])
@pytest.mark.parametrize("name", ["hier_noise_lastback", "flat_softplus_white"])
def test_points_forward_matches_golden_of_the_real_methods(pkg, name, backend):
    from _integrate_cases import check_points_golden
    from _util import build_generator
    from oracle import cips3d_oracle as O
    G = build_generator(DEV, O.synthetic_state_dict(O.generator_template(), seed=77, sigma_bias=0.3))
    check_points_golden(name, pkg, G, DEV, backend)
