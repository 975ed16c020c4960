"""Image metrics of the reference's test loop, computed on the device through the C ABI.

  img2mse, mse2psnr     utils/run_nerf_raybased_helpers.py:19-20
  ssim                  utils/ssim_torch.py:10-94 through the wrapper of main.py:46 ([C,H,W] in, scalar out)
  image_errors          the per-frame stage of render_path, main.py:330-335: abs error map, PSNR and SSIM of a
                        batch of frames in two kernel launches
  FLIP, flip_map        utils/flip_loss.py:56-140 as called from main.py:370-379 (LDR-FLIP of the rescaled stacks)
LPIPS (a pretrained network) is outside the hot-path scope.
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


# ------------------------------------------------------------------------------------------------ FLIP
def _flip_filters(pixels_per_degree):
    """The five filter kernels of LDR-FLIP as fp32 arrays, built like the reference builds them (numpy doubles, cast
    to fp32 last): contrast-sensitivity Gaussians A / RG / BY (utils/flip_loss.py:143-190) and the edge / point
    detectors (first / second x-derivative of a Gaussian, positive and negative lobes normalised separately, :263-290).
    Returns (taps [3*n_csf + 2*n_feat] fp32, r_csf, r_feat)."""
    ppd = float(pixels_per_degree)
    csf = {"A": (1., 0.0047, 0., 1e-5), "RG": (1., 0.0053, 0., 1e-5), "BY": (34.1, 0.04, 13.5, 0.025)}
    b_max = max(max(v[1], v[3]) for v in csf.values())
    r = int(np.ceil(3 * np.sqrt(b_max / (2 * np.pi**2)) * ppd))
    x, y = np.meshgrid(range(-r, r + 1), range(-r, r + 1))
    z = (x / ppd)**2 + (y / ppd)**2
    taps = []
    for name in ("A", "RG", "BY"):
        a1, b1, a2, b2 = csf[name]
        g = a1 * np.sqrt(np.pi / b1) * np.exp(-np.pi**2 * z / b1) + a2 * np.sqrt(np.pi / b2) * np.exp(-np.pi**2 * z / b2)
        taps.append(torch.Tensor(g / np.sum(g)).numpy().ravel())
    sd = 0.5 * 0.082 * ppd
    rf = int(np.ceil(3 * sd))
    x, y = np.meshgrid(range(-rf, rf + 1), range(-rf, rf + 1))
    g = np.exp(-(x**2 + y**2) / (2 * sd * sd))
    for gx in (np.multiply(-x, g), np.multiply(x**2 / (sd * sd) - 1, g)):
        neg, pos = -np.sum(gx[gx < 0]), np.sum(gx[gx > 0])
        t = torch.Tensor(gx)
        taps.append(torch.where(t < 0, t / neg, t / pos).numpy().ravel())
    return np.ascontiguousarray(np.concatenate(taps), dtype=np.float32), r, rf


def _flip_consts(qc, qf, pc, pt):
    """Colour-space constants in fp32 the way the reference obtains them (utils/flip_loss.py:293-306): the sRGB->XYZ
    matrix, its torch.inverse, the D65 illuminant A @ 1, and cmax = HyAB(green, blue)^qc after the Hunt adjustment."""
    A = torch.Tensor([[10135552 / 24577794, 8788810 / 24577794, 4435075 / 24577794],
                      [2613072 / 12288897, 8788810 / 12288897, 887015 / 12288897],
                      [1425312 / 73733382, 8788810 / 73733382, 70074185 / 73733382]])
    Ainv = torch.inverse(A)
    illum = torch.matmul(A, torch.ones(3, 1))[:, 0]

    def hunt_lab(rgb):
        t = torch.matmul(A, torch.tensor(rgb).reshape(3, 1))[:, 0] / illum
        delta = 6 / 29
        f = torch.where(t > 0.00885, torch.pow(t, 1 / 3), t / (3 * delta * delta) + 4 / 29)
        L = 116 * f[1] - 16
        return L, (0.01 * L) * (500 * (f[0] - f[1])), (0.01 * L) * (200 * (f[1] - f[2]))
    g, b = hunt_lab([0., 1., 0.]), hunt_lab([0., 0., 1.])
    hy = (g[0] - b[0]).abs() + torch.sqrt((g[1] - b[1])**2 + (g[2] - b[2])**2)
    cmax = torch.pow(hy, qc).item()
    c = np.concatenate([A.numpy().ravel(), Ainv.numpy().ravel(), illum.numpy(), [cmax, qc, qf, pc, pt]])
    return np.ascontiguousarray(c, dtype=np.float32)


_FLIP_CACHE = {}


def flip_map(test, reference, pixels_per_degree=None, scale=(1., 0.), want_map=True):
    """LDR-FLIP of image stacks in the renderer's layout: test, reference [N, H, W, 3] (or [H, W, 3]); every value is
    mapped v -> v * scale[0] + scale[1] first (or ((s_t, o_t), (s_r, o_r)) per stack).  Returns (map [N, H, W] | None,
    per-image mean [N]) as device tensors."""
    test, reference = _hwc_pair(test, reference, "flip")
    single = test.dim() == 3
    if single:
        test, reference = test[None], reference[None]
    if test.dim() != 4 or test.shape[-1] != 3:
        raise ValueError(f"images must be [N, H, W, 3], got {tuple(test.shape)}")
    if pixels_per_degree is None:
        pixels_per_degree = 0.7 * (3840 / 0.7) * (np.pi / 180)
    N, H, W, _ = test.shape
    dev = test.device
    key = (float(pixels_per_degree), dev.index)
    if key not in _FLIP_CACHE:
        taps, r, rf = _flip_filters(pixels_per_degree)
        _FLIP_CACHE[key] = (torch.from_numpy(taps).to(dev), r, rf, _flip_consts(0.7, 0.5, 0.4, 0.95))
    taps, r, rf, consts = _FLIP_CACHE[key]
    (s_t, o_t), (s_r, o_r) = (scale, scale) if not isinstance(scale[0], (tuple, list)) else scale
    ws = torch.empty((N, 2, 3, H, W), dtype=torch.float32, device=dev)
    out = torch.empty((N, H, W), dtype=torch.float32, device=dev) if want_map else None
    sums = torch.empty(max(N, 1), dtype=torch.float64, device=dev)
    with torch.cuda.device(dev):
        _lib.call("r2l_flip", N, H, W, _lib.ptr(test), _lib.ptr(reference), H * W * 3, float(s_t), float(o_t), float(s_r),
                  float(o_r), ctypes.c_void_p(consts.ctypes.data), r, rf, _lib.ptr(taps), _lib.ptr(ws), _lib.ptr(out),
                  _lib.ptr(sums), _lib.stream_ptr(dev))
    mean = (sums[:N] / float(H * W)).to(torch.float32)
    if single:
        return (out[0] if out is not None else None), mean[0]
    return out, mean


class FLIP(torch.nn.Module):
    """utils/flip_loss.py:56-140 on the device: same attributes, `compute_flip(reference, test, pixels_per_degree)` on
    [N, 3, H, W] tensors -> [N, 1, H, W], `forward(outputs, targets)` -> mean.  (A permuted view of an [N, H, W, 3]
    stack — what main.py:362-363 passes — is used in place.)"""

    def __init__(self):
        super().__init__()
        self.monitor_distance = 0.7
        self.monitor_width = 0.7
        self.monitor_resolution_x = 3840
        self.pixels_per_degree = self.monitor_distance * (self.monitor_resolution_x / self.monitor_width) * (np.pi / 180)
        self.qc, self.qf, self.pc, self.pt = 0.7, 0.5, 0.4, 0.95

    def compute_flip(self, reference, test, pixels_per_degree):
        if reference.dim() != 4 or reference.shape[1] != 3:
            raise ValueError(f"compute_flip expects [N, 3, H, W] images, got {tuple(reference.shape)}")
        m, _ = flip_map(test.permute(0, 2, 3, 1), reference.permute(0, 2, 3, 1), pixels_per_degree)
        return m[:, None]

    def forward(self, outputs, targets):
        return torch.mean(self.compute_flip(targets, outputs, self.pixels_per_degree))


class FLIPLoss():
    """utils/flip_loss.py:47-54."""

    def __init__(self):
        self.model = FLIP()

    def __call__(self, outputs, targets):
        return self.model.forward(outputs, targets)
