"""Image export on the GPU: c3d_image_to_u8 bit-exact against the oracle (which is pinned to
torchvision / PIL in tests/test_image_export_cpu.py) and gen_images with the real generator."""
import numpy as np
import pytest
import torch

from _util import build_generator
from oracle import cips3d_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.fixture(scope="module")
def pkg():
    import cips3d_b200
    cips3d_b200._lib.load()
    return cips3d_b200


def _images(shape, seed, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    x = torch.tanh(torch.randn(*shape, generator=g) * 1.5) * scale
    k = torch.arange(0, 256, dtype=torch.float32)
    ties = torch.cat([k / 255 * 2 - 1, (k + 0.5) / 255 * 2 - 1, torch.tensor([-1.0, 1.0, 0.0, -0.0])]) * scale
    x.view(-1)[:ties.numel()] = ties[:x.numel()]
    return x


@pytest.mark.parametrize("layout", ["nchw", "view_of_nhwc"])
@pytest.mark.parametrize("mode", ["save_image", "tensor_to_pil", "to_pil"])
@pytest.mark.parametrize("shape", [(4, 3, 256, 256), (2, 3, 37, 53), (3, 1, 64, 64), (1, 4, 32, 32)])
def test_image_to_u8_bit_exact(pkg, mode, shape, layout):
    x = _images(shape, seed=sum(shape), scale=1.0 if mode == "to_pil" else 1.25)
    want = O.image_to_u8(x, mode)
    xd = x.to(DEV)
    if layout == "view_of_nhwc":
        xd = xd.permute(0, 2, 3, 1).contiguous().permute(0, 3, 1, 2)
    got = pkg.ops.image_to_u8(xd, mode=mode)
    assert got.dtype == torch.uint8 and got.device == xd.device
    assert torch.equal(got.cpu(), want)


def test_image_to_u8_matches_torch_cuda_chain(pkg):
    """The five torch ops of make_grid.norm_ip + save_image on the same GPU tensor (what the reference executes)."""
    x = _images((8, 3, 128, 128), seed=1, scale=1.25).to(DEV)
    t = x.clone()
    t.clamp_(min=-1, max=1)
    t.sub_(-1).div_(max(1 - (-1), 1e-5))
    want = t.mul(255).add_(0.5).clamp_(0, 255).permute(0, 2, 3, 1).to(torch.uint8)
    assert torch.equal(pkg.ops.image_to_u8(x), want)


def test_gen_images_real_generator(pkg, tmp_path):
    """gen_images.py:30-73 with the fused generator at r32: the files equal torchvision's save_image of the same batches."""
    tv = pytest.importorskip("torchvision.utils")
    sd = O.synthetic_state_dict(O.generator_template(), seed=77, sigma_bias=0.3)
    G = build_generator(DEV, sd)
    seen = []
    fwd = G.forward

    def spy(*a, **k):
        out = fwd(*a, **k)
        seen.append(out[0].detach().clone())
        return out
    G.forward = spy
    kw = dict(O.G_KWARGS)
    kw["num_steps"] = 6
    torch.manual_seed(3)
    n = pkg.inference.gen_images(0, 1, G, kw, str(tmp_path / "fake"), num_imgs=5, img_size=32, batch_size=2, ext="png")
    assert n == 6 and len(seen) == 3
    from PIL import Image
    for idx_b, imgs in enumerate(seen):
        for idx_i, img in enumerate(imgs):
            name = f"{idx_b * 2 + idx_i:0>5}.png"
            ref = tmp_path / ("ref_" + name)
            tv.save_image(img, str(ref), normalize=True, value_range=(-1, 1))
            assert np.array_equal(np.array(Image.open(tmp_path / "fake" / name)), np.array(Image.open(ref))), name
