"""r2l_b200: B200-native (sm_100a) per-ray rendering hot path of R2L / NeRF.

Mirrors the reference's call surface:
  run_nerf_raybased_helpers  <- utils/run_nerf_raybased_helpers.py (get_rays, ndc_rays, Embedder, get_embedder,
                                raw2outputs, sample_pdf)
  nerf_raybased              <- model/nerf_raybased.py (NeRF, ResMLP, NeRF_v3_2, PointSampler, PositionalEmbedder)
  render                     <- main.py / utils/create_data.py render glue (render, render_rays, batchify_rays, ...)
  metrics                    <- img2mse / mse2psnr (helpers:19-20), utils/ssim_torch.py: render_path's PSNR / SSIM stage
  create_data                <- utils/create_data.py `--create_data rand` (pseudo-data shards, sharded over ranks)
All compute goes through the C-ABI CUDA library (include/r2l_b200.h); there is no CPU fallback.
"""
from . import _lib
from . import run_nerf_raybased_helpers
from . import nerf_raybased
from . import render
from . import sharding
from . import create_data
from . import compat
from . import metrics
from . import synthetic
from . import dropin
from .run_nerf_raybased_helpers import (get_rays, ndc_rays, Embedder, get_embedder, raw2outputs, sample_pdf,
                                        normalize_dirs, merge_sorted)
from .nerf_raybased import NeRF, ResMLP, NeRF_v3_2, PointSampler, PositionalEmbedder, LazyEmbedding, LazyPoints, get_activation
from .render import render_rays, render_rays_create_data, batchify_rays, batchify, run_network, render_r2l, render_path, GraphedR2L
from .render import render as render_image

__version__ = "0.1.0"
