"""Loader for the REAL reference (unmodified files of MingSun-Tse/Efficient-NeRF) — TEST / BASELINE INFRASTRUCTURE ONLY.

The reference's sources are never part of this repo; `baseline/stage_reference.py` copies them byte for byte into the
git-ignored `baseline/_ref/` (which travels to the GPU box) and pins their hashes in `baseline/MANIFEST.sha256`.
This module imports them from there (or from /root/reference in the build container):

  * model/nerf_raybased.py and utils/run_nerf_raybased_helpers.py are imported AS THEY ARE, by file path, under private
    module names (so they never collide with the product's stand-ins for `model.*` / `utils.*`);
  * main.py cannot be imported (`option.py` parses argv at import; smilelogging / imageio / lpips are absent), so the
    six render functions (main.py:51-186, 556-756) are AST-extracted and exec'd verbatim with the names they use —
    the recipe of SURVEY.md Appendix C, the same one oracle/make_golden.py pins the goldens with.

Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may use this module (like everything under oracle/).
"""
import ast
import importlib.util
import os
import sys
from types import SimpleNamespace

import numpy as np
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
MAIN_FUNCS = ("batchify", "run_network", "batchify_rays", "render", "raw2outputs", "render_rays")


def reference_root(required=True):
    spec = importlib.util.spec_from_file_location("_r2l_stage_reference", os.path.join(ROOT, "baseline", "stage_reference.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.staged_root(required)


def _import_path(name, path):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


_CACHE = {}


def load(root=None, device=None):
    """-> namespace(root, M = model/nerf_raybased.py, Hh = utils/run_nerf_raybased_helpers.py, main = dict of the six
    render functions of main.py, device).  `device`: the module-global `device` the reference computes once at import
    (model:10, helpers:13); default = the reference's own choice (cuda if available)."""
    root = root or reference_root()
    key = (root, str(device))
    if key in _CACHE:
        return _CACHE[key]
    tag = f"_r2l_ref_{len(_CACHE)}"
    M = _import_path(tag + "_model_nerf_raybased", os.path.join(root, "model", "nerf_raybased.py"))
    Hh = _import_path(tag + "_utils_helpers", os.path.join(root, "utils", "run_nerf_raybased_helpers.py"))
    torch.autograd.set_detect_anomaly(False)     # both modules switch anomaly mode on at import (model:4, helpers:8)
    if device is not None:
        M.device = Hh.device = torch.device(device)
    dev = M.device
    with open(os.path.join(root, "main.py")) as f:
        src = f.read()
    ns = dict(torch=torch, np=np, F=F, device=dev, to_tensor=Hh.to_tensor, get_rays=Hh.get_rays, ndc_rays=Hh.ndc_rays,
              sample_pdf=Hh.sample_pdf, DEBUG=False)
    for node in ast.parse(src).body:
        if isinstance(node, ast.FunctionDef) and node.name in MAIN_FUNCS:
            exec(compile(ast.Module([node], []), os.path.join(root, "main.py"), "exec"), ns)
    missing = [n for n in MAIN_FUNCS if n not in ns]
    if missing:
        raise RuntimeError(f"main.py lacks {missing}")
    out = SimpleNamespace(root=root, M=M, Hh=Hh, main=ns, device=dev)
    _CACHE[key] = out
    return out


def r2l_args(netdepth=88, netwidth=256):
    """The options NeRF_v3_2 reads for README.md:51's command (option.py defaults for the rest)."""
    return SimpleNamespace(netdepth=netdepth, netwidth=netwidth, layerwise_netwidths='', act='relu', linear_tail=False,
                           use_residual=True,
                           trial=SimpleNamespace(inact='relu', outact='none', body_arch='resmlp', n_block=-1,
                                                 res_scale=1., n_learnable=2))


def build_r2l(ref, seed=0, H=400, W=400, focal=555.5555155968841, n_sample=16, near=2., far=6.):
    """create_nerf's R2L branch (main.py:413-422) + main.py:1017: seeded model, PositionalEmbedder, PointSampler."""
    torch.manual_seed(seed)
    model = ref.M.NeRF_v3_2(r2l_args(), 3 * n_sample * 21, 3).to(ref.device).eval()
    pe = ref.M.PositionalEmbedder(L=10)
    ps = ref.M.PointSampler(H, W, focal, n_sample, near, far)
    return model, pe, ps


def build_nerf(ref, seed=0):
    """create_nerf's NeRF branch (main.py:424-453): coarse then fine NeRF (seeded), embedders, network_query_fn."""
    torch.manual_seed(seed)
    embed_fn, input_ch = ref.Hh.get_embedder(10, 0)
    embeddirs_fn, input_ch_views = ref.Hh.get_embedder(4, 0)
    coarse = ref.M.NeRF(D=8, W=256, input_ch=input_ch, output_ch=5, skips=[4], input_ch_views=input_ch_views,
                        use_viewdirs=True).to(ref.device).eval()
    fine = ref.M.NeRF(D=8, W=256, input_ch=input_ch, output_ch=5, skips=[4], input_ch_views=input_ch_views,
                      use_viewdirs=True).to(ref.device).eval()
    run_network = ref.main["run_network"]

    def network_query_fn(inputs, viewdirs, network_fn):
        return run_network(inputs, viewdirs, network_fn, embed_fn=embed_fn, embeddirs_fn=embeddirs_fn, netchunk=1024 * 64)

    kw = dict(network_query_fn=network_query_fn, perturb=0., N_importance=128, network_fine=fine, N_samples=64,
              network_fn=coarse, use_viewdirs=True, white_bkgd=True, raw_noise_std=0., ndc=False, near=2., far=6.)
    return coarse, fine, kw
