"""ctypes binding of libr2l_b200.so (the C ABI declared in include/r2l_b200.h).

There is no CPU fallback: if the library is missing or a call fails, a RuntimeError is raised.
"""
import ctypes
import os
import threading

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libr2l_b200.so")

_c_f32p = ctypes.c_void_p  # device float*
_c_ll = ctypes.c_longlong
_c_int = ctypes.c_int
_c_dbl = ctypes.c_double
_c_vp = ctypes.c_void_p

# name -> argtypes (restype is int unless listed in _RESTYPES); must match include/r2l_b200.h
SIGNATURES = {
    "r2l_last_error": [],
    "r2l_abi_version": [],
    "r2l_kernel_launches": [],
    "r2l_get_rays": [_c_int, _c_int, _c_dbl, _c_f32p, _c_f32p, _c_f32p, _c_vp],
    "r2l_ndc_rays": [_c_ll, _c_int, _c_int, _c_dbl, _c_dbl, _c_f32p, _c_f32p, _c_f32p, _c_f32p, _c_vp],
    "r2l_normalize_dirs": [_c_ll, _c_f32p, _c_ll, _c_f32p, _c_vp],
    "r2l_z_vals": [_c_ll, _c_int, _c_f32p, _c_f32p, _c_ll, _c_f32p, _c_int, _c_f32p, _c_f32p, _c_vp],
    "r2l_points_from_rays": [_c_ll, _c_int, _c_f32p, _c_ll, _c_f32p, _c_ll, _c_f32p, _c_ll, _c_f32p, _c_vp],
    "r2l_point_sample": [_c_int, _c_int, _c_dbl, _c_f32p, _c_f32p, _c_int, _c_f32p, _c_vp],
    "r2l_point_sample_batch": [_c_int, _c_int, _c_int, _c_dbl, _c_f32p, _c_f32p, _c_int, _c_f32p, _c_vp],
    "r2l_plucker": [_c_ll, _c_f32p, _c_ll, _c_f32p, _c_ll, _c_f32p, _c_vp],
    "r2l_embed": [_c_ll, _c_int, _c_int, _c_int, _c_int, _c_f32p, _c_f32p, _c_vp],
    "r2l_raw2outputs": [_c_ll, _c_int, _c_f32p, _c_f32p, _c_f32p, _c_ll, _c_f32p, _c_int, _c_f32p, _c_f32p,
                        _c_f32p, _c_f32p, _c_f32p, _c_vp],
    "r2l_sample_pdf": [_c_ll, _c_int, _c_int, _c_f32p, _c_ll, _c_f32p, _c_ll, _c_f32p, _c_int, _c_f32p, _c_vp,
                       _c_vp],
    "r2l_hier_sample": [_c_ll, _c_int, _c_int, _c_f32p, _c_f32p, _c_f32p, _c_int, _c_f32p, _c_f32p, _c_f32p, _c_vp,
                        _c_vp],
    "r2l_merge_sorted": [_c_ll, _c_int, _c_int, _c_f32p, _c_f32p, _c_f32p, _c_f32p, _c_vp],
    "r2l_image_error": [_c_int, _c_ll, _c_f32p, _c_f32p, _c_f32p, _c_vp, _c_vp],
    "r2l_ssim": [_c_int, _c_int, _c_int, _c_f32p, _c_f32p, _c_ll, _c_vp, _c_vp, _c_vp],
    "r2l_flip": [_c_int, _c_int, _c_int, _c_f32p, _c_f32p, _c_ll, _c_dbl, _c_dbl, _c_dbl, _c_dbl, _c_vp, _c_int, _c_int,
                 _c_f32p, _c_f32p, _c_f32p, _c_vp, _c_vp],
    "r2l_linear_fp32": [_c_ll, _c_int, _c_int, _c_f32p, _c_ll, _c_f32p, _c_ll, _c_f32p, _c_f32p, _c_ll, _c_int,
                        _c_f32p, _c_ll, _c_vp],
    "r2l_nerf_create": [ctypes.POINTER(_c_vp), _c_int, ctypes.POINTER(_c_vp), ctypes.POINTER(_c_vp), _c_f32p,
                        _c_f32p, _c_f32p, _c_f32p, _c_f32p, _c_f32p, _c_f32p, _c_f32p, _c_vp],
    "r2l_nerf_forward": [_c_vp, _c_ll, _c_int, _c_f32p, _c_ll, _c_f32p, _c_ll, _c_f32p, _c_ll, _c_f32p, _c_f32p,
                         _c_vp],
    "r2l_nerf_render": [_c_vp, _c_ll, _c_int, _c_f32p, _c_ll, _c_f32p, _c_ll, _c_f32p, _c_ll, _c_f32p, _c_int, _c_f32p,
                        _c_f32p, _c_f32p, _c_f32p, _c_f32p, _c_vp],
    "r2l_nerf_render_mode": [_c_vp, _c_int],
    "r2l_nerf_far_fixup": [_c_vp, _c_int, _c_dbl, _c_dbl],
    "r2l_nerf_far_count": [_c_vp, ctypes.POINTER(_c_ll), _c_vp],
    "r2l_nerf_forward_embedded": [_c_vp, _c_ll, _c_f32p, _c_ll, _c_f32p, _c_vp],
    "r2l_nerf_profile": [_c_vp, _c_ll, _c_int, _c_f32p, _c_ll, _c_f32p, _c_ll, _c_f32p, _c_ll, _c_f32p, _c_f32p,
                         _c_vp, _c_vp],
    "r2l_resmlp_create": [ctypes.POINTER(_c_vp), _c_int, _c_int, _c_int, _c_f32p, _c_f32p, ctypes.POINTER(_c_vp),
                          ctypes.POINTER(_c_vp), ctypes.POINTER(_c_vp), ctypes.POINTER(_c_vp), _c_dbl, _c_f32p,
                          _c_f32p, _c_int, _c_int, _c_vp],
    "r2l_resmlp_forward": [_c_vp, _c_ll, _c_f32p, _c_ll, _c_f32p, _c_vp],
    "r2l_resmlp_forward_embedded": [_c_vp, _c_ll, _c_f32p, _c_ll, _c_f32p, _c_vp],
    "r2l_resmlp_render": [_c_vp, _c_int, _c_int, _c_int, _c_dbl, _c_f32p, _c_f32p, _c_int, _c_ll, _c_ll, _c_f32p,
                          ctypes.POINTER(_c_vp), _c_int, _c_ll, _c_vp],
    "r2l_resmlp_forward_gather": [_c_vp, _c_ll, _c_f32p, _c_ll, ctypes.POINTER(_c_vp), _c_int, _c_ll, _c_vp],
    "r2l_resmlp_debug_head": [_c_vp, _c_ll, _c_f32p, _c_ll, _c_f32p, _c_f32p, _c_f32p, _c_vp, _c_vp],
    "r2l_resmlp_profile": [_c_vp, _c_ll, _c_f32p, _c_ll, _c_f32p, _c_vp, _c_vp],
    "r2l_mlp_debug_wstream": [_c_vp, _c_vp, ctypes.c_ulonglong, ctypes.POINTER(ctypes.c_ulonglong)],
    "r2l_mlp_destroy": [_c_vp],
    "r2l_mlp_status": [_c_vp, _c_vp],
    "r2l_tc_gemm_probe": [_c_int, _c_int, _c_int, _c_f32p, _c_f32p, _c_f32p, _c_int, _c_vp],
    "r2l_tc_gemm_probe_pair": [_c_int, _c_int, _c_int, _c_f32p, _c_f32p, _c_f32p, _c_vp],
}
_RESTYPES = {"r2l_last_error": ctypes.c_char_p, "r2l_kernel_launches": ctypes.c_longlong}

ABI_VERSION = 7   # r2l_abi_version() of the library this binding was written for (include/r2l_b200.h)

_lock = threading.Lock()
_lib = None
launch_count = 0  # number of C-ABI compute calls issued (bench.py reports it)


def load():
    """Load the shared library (once).  Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} not found: build it with `python efficient-nerf_b200/build.py` "
                "(or __graft_entry__.build()).  There is no CPU fallback.")
        lib = ctypes.CDLL(LIB_PATH)
        for name, argtypes in SIGNATURES.items():
            fn = getattr(lib, name)  # AttributeError if the symbol is not exported
            fn.argtypes = argtypes
            fn.restype = _RESTYPES.get(name, ctypes.c_int)
        lib.r2l_abi_version.restype = ctypes.c_int
        if lib.r2l_abi_version() != ABI_VERSION:
            raise RuntimeError(f"{LIB_PATH} has ABI version {lib.r2l_abi_version()}, this package needs {ABI_VERSION}: "
                               "rebuild it with `python efficient-nerf_b200/build.py`")
        _lib = lib
    return _lib


def check(rc, what=""):
    if rc != 0:
        msg = load().r2l_last_error()
        raise RuntimeError(f"r2l_b200 {what} failed (code {rc}): {msg.decode() if msg else '?'}")


def call(name, *args):
    global launch_count
    lib = load()
    launch_count += 1
    check(getattr(lib, name)(*args), name)


def kernel_launches():
    """CUDA kernels launched by the library so far (counted inside the library, not ABI calls)."""
    return int(load().r2l_kernel_launches())


def require_cuda():
    if not torch.cuda.is_available():
        raise RuntimeError("r2l_b200 needs a CUDA device (sm_100a); there is no CPU fallback")


def stream_ptr(device=None):
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def ptr(t):
    """Device pointer of a tensor (None -> NULL)."""
    if t is None:
        return ctypes.c_void_p(0)
    return ctypes.c_void_p(t.data_ptr())


def ptr_array(tensors):
    arr = (ctypes.c_void_p * len(tensors))()
    for i, t in enumerate(tensors):
        arr[i] = t.data_ptr()
    return arr


def as_f32_cuda(x, device=None, name="tensor"):
    """Contiguous fp32 CUDA tensor holding x (numpy arrays / CPU tensors are uploaded)."""
    require_cuda()
    if not isinstance(x, torch.Tensor):
        x = torch.as_tensor(x)
    if x.requires_grad and torch.is_grad_enabled():
        raise RuntimeError(f"{name}: the r2l_b200 kernels are inference-only; wrap the call in torch.no_grad() "
                           "or detach the input (training stays on the reference path)")
    if device is None:
        device = x.device if x.is_cuda else torch.device("cuda", torch.cuda.current_device())
    x = x.detach()
    if x.dtype != torch.float32 or x.device != device:
        x = x.to(device=device, dtype=torch.float32)
    return x.contiguous()
