"""Inference / evaluation wrappers that leave the GPU as uint8.

Mirrors, with the reference's names, arguments and file naming:
  gen_images(rank, world_size, generator, G_kwargs, fake_dir, num_imgs, img_size, batch_size)
      exp/cips3d/scripts/gen_images.py:30-73 -- the FID-evaluation image dump (2 048 - 50 000 images per evaluation)
  sample_images(rank, world_size, generator, G_kwargs, fake_dir, num_imgs, img_size)
      exp/cips3d/scripts/sample_images.py:30-83 -- one frontal-ish image per random seed, named by the seed
  to_pil(frame)            exp/comm/comm_utils.py:21-24
  tensor_to_PIL(img)       exp/cips3d/models/st_web.py:44-46
The reference clamps / rescales / rounds / permutes the fp32 batch with five torch ops, copies 4 bytes per sample to the
host image by image and encodes there.  Here the generator's output goes through ONE native kernel
(c3d_image_to_u8: the same fp32 arithmetic in the same order, so the bytes are identical), the batch leaves the device
as 1 byte per sample into pinned memory on a copy stream, and the host encodes batch k while the GPU renders batch k+1.
The files are written by the same PIL call torchvision's save_image ends in, so equal images give equal files."""
import copy
import math
import os

import numpy as np
import torch

from . import ops


def images_to_uint8(imgs, mode="save_image", value_range=(-1, 1)):
    """(B, 3, H, W) generator output -> (B, H, W, 3) uint8 on the same device (ops.image_to_u8)."""
    return ops.image_to_u8(imgs, mode=mode, value_range=value_range)


def _pil():
    from PIL import Image
    return Image


def to_pil(frame):
    """comm_utils.py:21-24: ((frame.squeeze() + 1) * 0.5) -> torchvision to_pil_image."""
    with torch.no_grad():
        u8 = ops.image_to_u8(frame.squeeze(), mode="to_pil")
    return _pil().fromarray(_squeeze_gray(u8.cpu().numpy()))


def tensor_to_PIL(img):
    """st_web.py:44-46."""
    with torch.no_grad():
        u8 = ops.image_to_u8(img.squeeze(), mode="tensor_to_pil")
    return _pil().fromarray(_squeeze_gray(u8.cpu().numpy()))


def _squeeze_gray(arr):
    return arr[:, :, 0] if arr.ndim == 3 and arr.shape[2] == 1 else arr


class _HostRing:
    """Two pinned host buffers + a copy stream: batch k is copied / encoded while batch k+1 is generated."""

    def __init__(self, device):
        self.cuda = torch.device(device).type == "cuda"
        self.stream = torch.cuda.Stream(device=device) if self.cuda else None
        self.bufs, self.events, self.k = [None, None], [None, None], 0

    def push(self, u8):
        """Start the device -> host copy of one uint8 batch; returns the slot to wait() on."""
        i = self.k & 1
        self.k += 1
        if self.bufs[i] is None or self.bufs[i].shape != u8.shape:
            self.bufs[i] = torch.empty(u8.shape, dtype=torch.uint8, pin_memory=self.cuda)
        if self.cuda:
            ready = torch.cuda.Event()
            ready.record()                                   # u8 is complete on the compute stream
            with torch.cuda.stream(self.stream):
                self.stream.wait_event(ready)
                self.bufs[i].copy_(u8, non_blocking=True)
                u8.record_stream(self.stream)
                self.events[i] = torch.cuda.Event()
                self.events[i].record()
        else:
            self.bufs[i].copy_(u8)
        return i

    def wait(self, i):
        if self.cuda:
            self.events[i].synchronize()
        return self.bufs[i].numpy()


def _synchronize():          # ddp_utils.d2_synchronize: a barrier when a process group exists
    if torch.distributed.is_available() and torch.distributed.is_initialized() and torch.distributed.get_world_size() > 1:
        torch.distributed.barrier()


def gen_images(rank, world_size, generator, G_kwargs, fake_dir, num_imgs, img_size, batch_size, ext="jpg", progress=False,
               forward_points=256 ** 2):
    """gen_images.py:30-73.  Every rank renders batch_size // world_size images per iteration and writes
    f"{fake_dir}/{idx_b * batch_size + idx_i * world_size + rank:0>5}.{ext}" (the reference's interleaving), each file
    through PIL exactly as save_image(img, path, normalize=True, value_range=(-1, 1)) does.  Returns the number of files
    this rank wrote."""
    if rank == 0:
        os.makedirs(fake_dir, exist_ok=True)
    _synchronize()
    metadata = copy.deepcopy(G_kwargs)
    batch_gpu = batch_size // world_size
    metadata['img_size'] = img_size
    metadata['batch_size'] = batch_gpu
    metadata['psi'] = 1
    generator.eval()
    Image = _pil()
    ring, pending, written = None, None, 0

    def flush(p):
        idx_b, slot = p
        arr = ring.wait(slot)
        for idx_i in range(arr.shape[0]):
            path = f"{fake_dir}/{idx_b * batch_size + idx_i * world_size + rank:0>5}.{ext}"
            Image.fromarray(_squeeze_gray(arr[idx_i])).save(path)
        return arr.shape[0]

    bar = None
    if progress and rank == 0:
        import tqdm
        bar = tqdm.tqdm(desc=f"Generating images at {img_size}x{img_size}", total=num_imgs)
    with torch.no_grad():
        for idx_b in range((num_imgs + batch_size - 1) // batch_size):
            if bar is not None:
                bar.update(batch_size)
            zs = generator.get_zs(metadata['batch_size'])
            generated_imgs = generator(zs, forward_points=forward_points, **metadata)[0]
            u8 = ops.image_to_u8(generated_imgs, mode="save_image", value_range=(-1, 1))
            if ring is None:
                ring = _HostRing(u8.device)
            slot = ring.push(u8)
            if pending is not None:
                written += flush(pending)          # encode batch k-1 on the host while batch k is in flight
            pending = (idx_b, slot)
        if pending is not None:
            written += flush(pending)
    if bar is not None:
        bar.close()
    _synchronize()
    return written


def sample_images(rank, world_size, generator, G_kwargs, fake_dir, num_imgs, img_size, ext="jpg", forward_points=256 ** 2, **kwargs):
    """sample_images.py:30-83: every rank draws a seed (`np.random.randint(0, 1e8)`), seeds torch with it, renders ONE image with
    the camera fixed at h_mean = pi/2 + 0.15 (h_stddev = v_stddev = 0, psi = 1) and writes f"{fake_dir}/{seed:0>10}.{ext}".
    Returns the seeds this rank used."""
    if rank == 0:
        os.makedirs(fake_dir, exist_ok=True)
    _synchronize()
    metadata = copy.deepcopy(G_kwargs)
    batch_size = world_size                       # batch_gpu = 1
    metadata['img_size'] = img_size
    metadata['batch_size'] = 1
    metadata['psi'] = 1
    metadata['h_stddev'] = 0
    metadata['v_stddev'] = 0
    metadata['h_mean'] = math.pi * 0.5 + 0.15
    generator.eval()
    Image = _pil()
    ring, pending, seeds = None, None, []

    def flush(p):
        seed, slot = p
        Image.fromarray(_squeeze_gray(ring.wait(slot)[0])).save(f"{fake_dir}/{seed:0>10}.{ext}")

    with torch.no_grad():
        for _ in range((num_imgs + batch_size - 1) // batch_size):
            seed = np.random.randint(0, 1e8)
            torch.manual_seed(seed)
            zs = generator.get_zs(metadata['batch_size'])
            generated_imgs = generator(zs, forward_points=forward_points, **metadata)[0]
            u8 = ops.image_to_u8(generated_imgs, mode="save_image", value_range=(-1, 1))
            if ring is None:
                ring = _HostRing(u8.device)
            slot = ring.push(u8)
            if pending is not None:
                flush(pending)
            pending = (seed, slot)
            seeds.append(int(seed))
        if pending is not None:
            flush(pending)
    _synchronize()
    return seeds
