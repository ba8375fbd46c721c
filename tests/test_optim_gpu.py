"""GPU parity of the optimiser tail (csrc/optim_ops.cu through the C-ABI) vs the oracle and vs torch.optim.Adam on
the same device.  Runs after test_gpu_parity.py (file order), so a failure here cannot hide the renderer results."""
import pytest
import torch

from _optim_cases import TOL, run_case

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.fixture(scope="module")
def pkg():
    import __graft_entry__ as g
    g.build()
    import cips3d_b200
    return cips3d_b200


@pytest.mark.parametrize("betas,max_norm", [((0.0, 0.999), 10.0), ((0.9, 0.99), 0.05), ((0.5, 0.999), None)])
def test_fused_adam_ema_matches_oracle(pkg, betas, max_norm):
    worst, _, _ = run_case(pkg, DEV, betas, max_norm)
    assert worst < TOL, worst


def test_fused_adam_many_tensors(pkg):
    worst, _, _ = run_case(pkg, DEV, (0.0, 0.999), 1.0, steps=2, many=2 * pkg.optim.OPT_MAX_TENSORS + 7, start_itr=0)
    assert worst < TOL, worst


def test_fused_adam_on_the_generator_matches_torch_adam(pkg):
    """All 119 parameter tensors of GeneratorNerfINR (11.3 M parameters), synthetic gradients, 3 steps, against
    torch.optim.Adam + clip_grad_norm_ + the reference EMA expression on the same GPU."""
    import copy
    from _util import build_generator
    from oracle import cips3d_oracle as O
    sd = O.synthetic_state_dict(O.generator_template(), seed=9)
    G1 = build_generator(DEV, sd).train()
    G2 = copy.deepcopy(G1)
    E1, E2 = copy.deepcopy(G1), copy.deepcopy(G2)
    o1 = pkg.FusedAdam(G1.parameters(), lr=2e-4, betas=(0.0, 0.999))
    o2 = torch.optim.Adam(G2.parameters(), lr=2e-4, betas=(0.0, 0.999))
    ema = pkg.EMA(G1, E1, decay=0.999, start_itr=1)
    g = torch.Generator(device=DEV).manual_seed(1)
    for it in range(3):
        for p1, p2 in zip(G1.parameters(), G2.parameters()):
            p1.grad = torch.randn(p1.shape, device=DEV, generator=g) * 0.1
            p2.grad = p1.grad.clone()
        n2 = torch.nn.utils.clip_grad_norm_(G2.parameters(), 10.0)
        o2.step()
        if it >= 1:
            s2, t2 = G2.state_dict(), E2.state_dict()
            for k in s2:
                t2[k].data.copy_(t2[k].data * 0.999 + s2[k].data * (1 - 0.999))
        n1 = o1.step(max_norm=10.0, ema=ema, itr=it, zero_grad=True)
        assert abs(float(n1) - float(n2)) < 1e-5 * float(n2)
    for (k, a), b in zip(G1.state_dict().items(), G2.state_dict().values()):
        assert (a - b).abs().max().item() <= TOL * b.abs().max().item() + 1e-9, k
    for (k, a), b in zip(E1.state_dict().items(), E2.state_dict().values()):
        assert (a - b).abs().max().item() <= TOL * b.abs().max().item() + 1e-9, k
