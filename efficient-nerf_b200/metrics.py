"""Image metrics of the reference's test loop, computed on the device through the C ABI.

  img2mse, mse2psnr     utils/run_nerf_raybased_helpers.py:19-20
  ssim                  utils/ssim_torch.py:10-94 through the wrapper of main.py:46 ([C,H,W] in, scalar out)
  image_errors          the per-frame stage of render_path, main.py:330-335: abs error map, PSNR and SSIM of a
                        batch of frames in two kernel launches
LPIPS (a pretrained network) and FLIP are outside the hot-path scope.
"""
import ctypes
from math import exp

import numpy as np
import torch

from . import _lib


def gaussian_taps(window_size=11, sigma=1.5):
    """The reference's 1-D window, computed the same way (ssim_torch.py:10-16): python-double exponentials stored in
    an fp32 tensor, normalised in fp32."""
    g = torch.Tensor([exp(-(x - window_size // 2)**2 / float(2 * sigma**2)) for x in range(window_size)])
    return g / g.sum()


def _hwc_pair(a, b, name):
    a = _lib.as_f32_cuda(a, name=name)
    b = _lib.as_f32_cuda(b, a.device, name)
    if a.shape != b.shape:
        raise ValueError(f"{name}: shapes differ: {tuple(a.shape)} vs {tuple(b.shape)}")
    return a, b


def image_errors(rgbs, gts, want_error_map=True, want_ssim=True):
    """rgbs, gts: [N, H, W, 3] (or one [H, W, 3] image).  Returns dict(errors [N,H,W,3] | None, mse [N], psnr [N],
    ssim [N] | None) — the per-frame quantities of main.py:330-335, as device tensors (fp32; sums are accumulated in
    double on the device)."""
    rgbs, gts = _hwc_pair(rgbs, gts, "image_errors")
    single = rgbs.dim() == 3
    if single:
        rgbs, gts = rgbs[None], gts[None]
    if rgbs.dim() != 4 or rgbs.shape[-1] != 3:
        raise ValueError(f"images must be [N, H, W, 3], got {tuple(rgbs.shape)}")
    N, H, W, _ = rgbs.shape
    dev = rgbs.device
    err = torch.empty_like(rgbs) if want_error_map else None
    sums = torch.empty(2, max(N, 1), dtype=torch.float64, device=dev)
    with torch.cuda.device(dev):
        _lib.call("r2l_image_error", N, H * W * 3, _lib.ptr(rgbs), _lib.ptr(gts), _lib.ptr(err), _lib.ptr(sums[0]),
                  _lib.stream_ptr(dev))
        if want_ssim:
            taps = np.ascontiguousarray(gaussian_taps().numpy(), dtype=np.float32)
            _lib.call("r2l_ssim", N, H, W, _lib.ptr(rgbs), _lib.ptr(gts), H * W * 3,
                      ctypes.c_void_p(taps.ctypes.data), _lib.ptr(sums[1]), _lib.stream_ptr(dev))
    mse = (sums[0, :N] / float(H * W * 3)).to(torch.float32)
    out = dict(errors=err, mse=mse, psnr=mse2psnr(mse),
               ssim=(sums[1, :N] / float(H * W * 3)).to(torch.float32) if want_ssim else None)
    if single:
        out = {k: (v[0] if v is not None else None) for k, v in out.items()}
    return out


def img2mse(x, y):
    """mean((x - y)^2) (helpers:19) over everything, as a 0-dim device tensor."""
    x, y = _hwc_pair(x, y, "img2mse")
    n = x.numel()
    s = torch.empty(1, dtype=torch.float64, device=x.device)
    with torch.cuda.device(x.device):
        _lib.call("r2l_image_error", 1, n, _lib.ptr(x), _lib.ptr(y), None, _lib.ptr(s), _lib.stream_ptr(x.device))
    return (s[0] / float(max(n, 1))).to(torch.float32)


def mse2psnr(x):
    """-10 log(x) / log(10) (helpers:20)."""
    x = x if isinstance(x, torch.Tensor) else torch.tensor(x, dtype=torch.float32)
    return -10. * torch.log(x) / torch.log(torch.tensor([10.], dtype=x.dtype, device=x.device))


def ssim(img, ref):
    """SSIM of one image pair in the reference's [C, H, W] layout (main.py:46); returns a 0-dim device tensor.
    A permuted view of an [H, W, 3] image (what render_path passes, main.py:333-335) is used in place."""
    if img.dim() != 3 or img.shape[0] != 3:
        raise ValueError(f"ssim expects [3, H, W] images, got {tuple(img.shape)}")
    a = img.permute(1, 2, 0)
    b = ref.permute(1, 2, 0)
    return image_errors(a, b, want_error_map=False)["ssim"]
