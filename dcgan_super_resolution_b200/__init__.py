"""dcgan_super_resolution_b200 -- B200-native (sm_100a) DCGAN super-resolution training step.

The product is ``libdcgansr.so`` (C ABI in ``include/dcgansr.h``; CUDA sources in ``csrc/``).  This
package is the thin Python host above it: a ctypes binding (``_lib``), the Torch7-shaped surface the
reference's ``train*.lua`` scripts use (``nn``), the reference's net graphs as data (``models``) and the
data-parallel plumbing (``parallel``), plus the Torch7 ``.t7`` checkpoint reader / writer (``t7``).  Nothing here computes on the CPU.
"""
from . import _lib, init, models, nn, parallel, t7  # noqa: F401
from ._lib import DcgansrError  # noqa: F401
from .nn import (Context, Sequential, assemble_patches, extract_patches, make_step_cfg, psnr, scale_bilinear, ssim, stage_batch, stage_patches,  # noqa: F401
                 stitch_overlap, train_step, train_step_staged)
