"""Checkpoint ingestion: load the reference's `.tar` checkpoints into the B200 modules.

The reference saves (main.py:1516-1542)
    {'global_step', 'best_psnr', 'best_psnr_step', 'network_fn_state_dict', 'optimizer_state_dict',
     ['network_fine_state_dict'] (NeRF with N_importance > 0), ['network_fn' = the pickled R2L module]}
and restores it in create_nerf (main.py:481-509): `ckpt['network_fn']` REPLACES the freshly built model, then
load_weights_v2 (helpers:362-382) loads the state dicts.  Unpickling `network_fn` needs the import paths
`model.nerf_raybased.{NeRF_v3_2, ResMLP}` and `utils.EmptyClass` (utils/__init__.py:1-2, the type smilelogging
gave to `args.trial`): install_reference_aliases() maps those paths onto this package, so the unpickled object IS a
B200 module (same attribute layout; the extra fields are filled in by __setstate__) and renders through the fused
kernels.  The DataParallel `module.` prefix is stripped like helpers:347-359 / :408-425 do.
"""
import sys
import types
from collections import OrderedDict

import torch

from . import nerf_raybased, run_nerf_raybased_helpers


class EmptyClass:
    """utils/__init__.py:1-2 — attribute bag used for the nested `args.trial` namespace in pickled checkpoints."""
    pass


def install_reference_aliases(force=False):
    """Make `model.nerf_raybased`, `utils.run_nerf_raybased_helpers` and `utils.EmptyClass` importable as aliases
    of this package (needed by torch.load of pickled reference modules).  Existing real modules are left alone
    unless `force`."""
    def ensure_pkg(name):
        mod = sys.modules.get(name)
        if mod is None or force:
            mod = types.ModuleType(name)
            mod.__path__ = []
            sys.modules[name] = mod
        return mod

    model_pkg = ensure_pkg("model")
    if force or "model.nerf_raybased" not in sys.modules:
        sys.modules["model.nerf_raybased"] = nerf_raybased
        model_pkg.nerf_raybased = nerf_raybased
    utils_pkg = ensure_pkg("utils")
    if force or "utils.run_nerf_raybased_helpers" not in sys.modules:
        sys.modules["utils.run_nerf_raybased_helpers"] = run_nerf_raybased_helpers
        utils_pkg.run_nerf_raybased_helpers = run_nerf_raybased_helpers
    if force or not hasattr(utils_pkg, "EmptyClass"):
        utils_pkg.EmptyClass = EmptyClass
        EmptyClass.__module__ = "utils"


def undataparallel(obj):
    """Remove the `module.` prefix / wrapper left by nn.DataParallel (helpers:408-425)."""
    if isinstance(obj, torch.nn.Module):
        return obj.module if hasattr(obj, 'module') else obj
    if isinstance(obj, dict):
        out = OrderedDict()
        for k, v in obj.items():
            out[k[7:] if k.startswith('module.') else k] = v
        return out
    raise NotImplementedError(type(obj))


def load_checkpoint(path, map_location="cpu"):
    """torch.load of a reference checkpoint (pickled modules allowed: weights_only=False, main.py:483)."""
    install_reference_aliases()
    return torch.load(path, map_location=map_location, weights_only=False)


def models_from_checkpoint(ckpt, model=None, model_fine=None, precision=None, device=None):
    """create_nerf's restore step (main.py:481-509): returns (model, model_fine) ready for rendering.
    `model` / `model_fine` are freshly constructed modules (may be None when the checkpoint pickles its own
    architecture, as R2L checkpoints do)."""
    if isinstance(ckpt, str):
        ckpt = load_checkpoint(ckpt)
    if 'network_fn' in ckpt:
        model = undataparallel(ckpt['network_fn'])
        if 'network_fine' in ckpt:
            model_fine = undataparallel(ckpt['network_fine'])
    if model is None:
        raise ValueError("the checkpoint holds no pickled 'network_fn'; pass a constructed model")
    model.load_state_dict(undataparallel(ckpt['network_fn_state_dict']))
    if model_fine is not None:
        model_fine.load_state_dict(undataparallel(ckpt['network_fine_state_dict']))
    for m in (model, model_fine):
        if m is None:
            continue
        if precision is not None:
            m.precision = nerf_raybased._check_precision(precision)
        if device is not None:
            m.to(device)
        m.eval()
    return model, model_fine
