"""cips3d_b200 -- B200-native (sm_100a) hot path of CIPS-3D behind the reference's class surface.

Layout:
  csrc/            hand-written CUDA kernels + the C-ABI (include/cips3d_b200.h)
  _lib.py          ctypes binding of libcips3d_b200.so (fails loudly when it is missing)
  ops.py           tensor-level entry points (volumetric renderer, CIPS MLP, D ops)
  pigan.py         pi-GAN surface: ImplicitGenerator3d, TALLSIREN, SPATIALSIRENBASELINE (piGAN_lib/generators, piGAN_lib/siren)
  optim.py         FusedAdam (+ gradient clipping, EMA, zero_grad in two launches), EMA (reference class surface)
  comm_utils.py    camera sampling / look-at matrices (host-side mirror, tiny)
  generator.py     GeneratorNerfINR, GeneratorNerfINR_freeze_NeRF, NeRFNetwork, CIPSNet, ...
  discriminator.py Discriminator_MultiScale(_Aux) and its layers
  inference.py     gen_images / to_pil / tensor_to_PIL: uint8 leaves the GPU (one native conversion kernel, pinned double buffer)
"""
from . import _lib  # noqa: F401
from . import ops  # noqa: F401
from . import optim  # noqa: F401
from .optim import EMA, FusedAdam  # noqa: F401
from . import pigan  # noqa: F401
from . import inference  # noqa: F401
from .pigan import ImplicitGenerator3d, SPATIALSIRENBASELINE, TALLSIREN  # noqa: F401
from .generator import (CIPSNet, FiLMLayer, GeneratorNerfINR, GeneratorNerfINR_freeze_NeRF,  # noqa: F401
                        MultiHeadMappingNetwork, NeRFNetwork, SinBlock, SinStyleMod, ToRGB)
from .discriminator import (Discriminator_MultiScale, Discriminator_MultiScale_Aux,  # noqa: F401
                            DiffAugment)

__version__ = "0.1.0"
