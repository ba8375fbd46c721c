"""Image export: the oracle's restatement pinned to the real torchvision / PIL path the reference's
scripts call, and the native kernel (CPU emulation of csrc/image_ops.cu) bit-exact against the oracle."""
import io

import numpy as np
import pytest
import torch

import _emu
from _emu import emulated
from oracle import cips3d_oracle as O


def _images(shape, seed, wide=False):
    g = torch.Generator().manual_seed(seed)
    x = torch.tanh(torch.randn(*shape, generator=g) * 1.5)
    if wide:        # values outside [-1, 1], exact grid points and rounding ties of the *255+0.5 step
        x = x * 1.3
        flat = x.view(-1)
        k = torch.arange(0, 256, dtype=torch.float32)
        ties = torch.cat([k / 255 * 2 - 1, (k + 0.5) / 255 * 2 - 1, torch.tensor([-1.0, 1.0, 0.0, -0.0, 1e-8, -1e-8])])
        flat[:ties.numel()] = ties[:flat.numel()]
    return x


@pytest.mark.parametrize("shape", [(3, 32, 32), (3, 17, 23), (1, 8, 8)])
def test_oracle_save_image_matches_torchvision_png_roundtrip(shape):
    """gen_images.py:64 -- save_image(img, path, normalize=True, value_range=(-1, 1)); PNG is lossless, so the decoded file
    is exactly the byte array torchvision handed to PIL."""
    tv = pytest.importorskip("torchvision.utils")
    Image = pytest.importorskip("PIL.Image")
    x = _images(shape, seed=shape[1], wide=True)
    buf = io.BytesIO()
    tv.save_image(x, buf, format="png", normalize=True, value_range=(-1, 1))
    got = np.array(Image.open(io.BytesIO(buf.getvalue())))
    want = O.image_to_u8(x, "save_image", (-1, 1)).numpy()
    if shape[0] == 1:           # make_grid repeats a single channel three times
        want = np.repeat(want, 3, axis=2)
    assert got.shape == want.shape and np.array_equal(got, want)


def test_oracle_to_pil_matches_torchvision_to_pil_image():
    """comm_utils.py:21-24 -- to_pil_image((frame.squeeze() + 1) * 0.5)."""
    F = pytest.importorskip("torchvision.transforms.functional")
    x = _images((1, 3, 24, 40), seed=5)
    got = np.array(F.to_pil_image((x.squeeze() + 1) * 0.5))
    assert np.array_equal(got, O.image_to_u8(x.squeeze(), "to_pil").numpy())


def test_oracle_tensor_to_pil_formula():
    """st_web.py:44-46, the one-liner itself."""
    x = _images((1, 3, 16, 16), seed=6, wide=True)
    img = x.squeeze() * 0.5 + 0.5
    want = img.mul(255).add_(0.5).clamp_(0, 255).permute(1, 2, 0).to('cpu', torch.uint8).numpy()
    assert np.array_equal(want, O.image_to_u8(x.squeeze(), "tensor_to_pil").numpy())


@pytest.mark.parametrize("layout", ["nchw", "view_of_nhwc"])
@pytest.mark.parametrize("mode", ["save_image", "tensor_to_pil", "to_pil"])
@pytest.mark.parametrize("shape", [(2, 3, 16, 16), (1, 3, 5, 7), (3, 1, 9, 4), (2, 4, 6, 6), (3, 8, 8)])
def test_emu_image_to_u8_bit_exact(mode, shape, layout):
    """The kernel (vectorised RGB path, the flat channels-last path the generator's output takes, and the generic paths),
    emulated on the CPU: every byte equal to the oracle's."""
    x = _images(shape, seed=sum(shape), wide=(mode != "to_pil"))
    if layout == "view_of_nhwc":        # what GeneratorNerfINR returns: .view(B, H, W, 3).permute(0, 3, 1, 2)
        if x.dim() == 3:
            pytest.skip("batched layout case")
        x = x.permute(0, 2, 3, 1).contiguous().permute(0, 3, 1, 2)
        assert not x.is_contiguous() or shape[1] == 1
    with emulated(async_mode=0) as pkg:
        got = pkg.ops.image_to_u8(x, mode=mode)
    want = O.image_to_u8(x, mode)
    assert got.dtype == torch.uint8 and got.shape == want.shape
    assert torch.equal(got, want)


def test_emu_image_to_u8_value_range_and_nan():
    x = _images((1, 3, 8, 8), seed=3, wide=True)
    with emulated(async_mode=0) as pkg:
        got = pkg.ops.image_to_u8(x, mode="save_image", value_range=(-0.5, 1.5))
        got2 = pkg.ops.image_to_u8(x, mode="save_image", value_range=(0.1, 0.9))     # not a power of two: true division
        xn = x.clone()
        xn[0, 0, 0, 0] = float("nan")
        nan_out = pkg.ops.image_to_u8(xn, mode="save_image")
        empty = pkg.ops.image_to_u8(torch.zeros(0, 3, 8, 8))
        with pytest.raises(Exception):
            pkg.ops.image_to_u8(torch.zeros(1, 5, 4, 4))
        with pytest.raises(ValueError):
            pkg.ops.image_to_u8(x, mode="jpeg")
    assert torch.equal(got, O.image_to_u8(x, "save_image", (-0.5, 1.5)))
    assert torch.equal(got2, O.image_to_u8(x, "save_image", (0.1, 0.9)))
    assert nan_out[0, 0, 0, 0].item() == 0
    assert empty.shape == (0, 8, 8, 3)


class _StubG:
    """Stands in for GeneratorNerfINR in gen_images (eval / get_zs / __call__ -> (imgs, pitch_yaw)); images are returned as
    the NCHW view of an NHWC buffer, like the real generator's."""

    def __init__(self):
        self.calls, self.images = [], []

    def eval(self):
        return self

    def get_zs(self, b):
        return {"z": torch.randn(b, 4)}

    def __call__(self, zs, forward_points=None, **kw):
        self.calls.append(dict(kw, forward_points=forward_points))
        b, r = zs["z"].shape[0], kw["img_size"]
        img = torch.tanh(torch.randn(b, r, r, 3) * 2).permute(0, 3, 1, 2)
        self.images.append(img.clone())
        return img, torch.zeros(b, 2)


@pytest.mark.parametrize("world_size,rank", [(1, 0), (2, 1)])
def test_gen_images_writes_the_files_save_image_writes(tmp_path, world_size, rank):
    """inference.gen_images vs gen_images.py:30-73: same file names (rank interleaving), same metadata passed to the
    generator, and every file byte-identical to torchvision's save_image(img, path, normalize=True, value_range=(-1, 1))."""
    tv = pytest.importorskip("torchvision.utils")
    import cips3d_b200.inference as inf
    G = _StubG()
    torch.manual_seed(7)
    if rank != 0:
        (tmp_path / "fake").mkdir()                    # rank 0 creates it in a real run (then the barrier)
    with emulated(async_mode=0):
        n = inf.gen_images(rank, world_size, G, {"fov": 12, "psi": 0.7}, str(tmp_path / "fake"), num_imgs=5, img_size=16,
                           batch_size=4)
    assert n == 2 * (4 // world_size)                  # ceil(5 / 4) batches of batch_size // world_size images
    assert all(c["img_size"] == 16 and c["psi"] == 1 and c["batch_size"] == 4 // world_size and c["fov"] == 12 and
               c["forward_points"] == 256 ** 2 for c in G.calls)
    for idx_b, imgs in enumerate(G.images):
        for idx_i, img in enumerate(imgs):
            name = f"{idx_b * 4 + idx_i * world_size + rank:0>5}.jpg"
            ref_path = tmp_path / ("ref_" + name)
            tv.save_image(img, str(ref_path), normalize=True, value_range=(-1, 1))
            assert (tmp_path / "fake" / name).read_bytes() == ref_path.read_bytes(), name


def test_to_pil_and_tensor_to_PIL_match_the_reference_helpers():
    F = pytest.importorskip("torchvision.transforms.functional")
    from PIL import Image
    import cips3d_b200.inference as inf
    x = _images((1, 3, 12, 20), seed=9)
    with emulated(async_mode=0):
        a, b = inf.to_pil(x), inf.tensor_to_PIL(x)
    assert np.array_equal(np.array(a), np.array(F.to_pil_image((x.squeeze() + 1) * 0.5)))
    img = x.squeeze() * 0.5 + 0.5
    want = Image.fromarray(img.mul(255).add_(0.5).clamp_(0, 255).permute(1, 2, 0).to('cpu', torch.uint8).numpy())
    assert np.array_equal(np.array(b), np.array(want)) and b.mode == want.mode == "RGB"


def test_sample_images_names_files_by_seed_and_fixes_the_camera(tmp_path):
    """sample_images.py:30-83: seed drawn from numpy, torch seeded with it, one image per seed, the reference's camera metadata."""
    tv = pytest.importorskip("torchvision.utils")
    import cips3d_b200.inference as inf
    G = _StubG()
    np.random.seed(123)
    with emulated(async_mode=0):
        seeds = inf.sample_images(0, 1, G, {"fov": 12, "h_stddev": 0.3, "psi": 0.5}, str(tmp_path / "s"), num_imgs=3, img_size=8)
    np.random.seed(123)
    assert seeds == [int(np.random.randint(0, 1e8)) for _ in range(3)]
    for c in G.calls:
        assert c["h_stddev"] == 0 and c["v_stddev"] == 0 and c["psi"] == 1 and c["batch_size"] == 1
        assert c["h_mean"] == pytest.approx(np.pi * 0.5 + 0.15) and c["forward_points"] == 256 ** 2
    for seed, img in zip(seeds, G.images):
        torch.manual_seed(seed)
        assert torch.equal(G.get_zs(1)["z"], torch.randn(1, 4, generator=torch.Generator().manual_seed(seed)))   # torch was seeded
        ref = tmp_path / f"ref_{seed}.jpg"
        tv.save_image(img.squeeze(), str(ref), normalize=True, value_range=(-1, 1))
        assert (tmp_path / "s" / f"{seed:0>10}.jpg").read_bytes() == ref.read_bytes()
