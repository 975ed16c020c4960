"""Checkpoint ingestion: a checkpoint written by the REAL reference classes
(oracle/make_golden_ckpt.py, reference format of main.py:1516-1542 incl. the pickled R2L module) loads into the
B200 modules through efficient_nerf_b200.compat — in a process that has no access to the reference sources."""
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

from conftest import t

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CKPT = os.path.join(ROOT, "tests", "golden", "ckpt_r2l_reference.tar")


def test_pickled_reference_module_unpickles_as_b200_module(E):
    ckpt = E.compat.load_checkpoint(CKPT)
    assert ckpt['global_step'] == 1234 and 'network_fn' in ckpt
    net, fine = E.compat.models_from_checkpoint(ckpt)
    assert fine is None and isinstance(net, E.NeRF_v3_2) and isinstance(net.body[0], E.ResMLP)
    assert net.input_dim == 252 and net.precision in ("fp16", "bf16", "fp32")
    assert net.supports_tensor_core_path()           # W256, resmlp, 4 points: covered by the fused kernel
    sd = ckpt['network_fn_state_dict']
    assert all(torch.equal(net.state_dict()[k], sd[k]) for k in sd)
    assert type(net.args.trial).__name__ == "EmptyClass" and net.args.trial.body_arch == 'resmlp'


def test_unpickling_needs_no_reference_sources():
    """fresh interpreter, reference NOT on sys.path: only the aliases make `model.nerf_raybased` resolvable"""
    code = ("import sys; sys.path.insert(0, %r); import efficient_nerf_b200 as E; "
            "assert not any('reference' in p for p in sys.path); "
            "c = E.compat.load_checkpoint(%r); print(type(c['network_fn']).__module__)") % (ROOT, CKPT)
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    assert "nerf_raybased" in r.stdout


def test_dataparallel_prefix_is_stripped(E, golden):
    g = golden("ckpt_reference_io")
    sd = {k[len("nerf_sd::"):]: t(g[k]) for k in g.files if k.startswith("nerf_sd::")}
    assert all(k.startswith("module.") for k in sd)
    net = E.NeRF(8, 256, 63, 27, 5, [4], True)
    coarse, _ = E.compat.models_from_checkpoint({'network_fn_state_dict': sd}, model=net)
    assert torch.equal(coarse.pts_linears[5].weight, sd['module.pts_linears.5.weight'])
    with pytest.raises(ValueError):
        E.compat.models_from_checkpoint({'network_fn_state_dict': sd})


@pytest.mark.gpu
def test_reference_checkpoint_renders_through_the_fused_kernels(E, golden):
    g = golden("ckpt_reference_io")
    net, _ = E.compat.models_from_checkpoint(E.compat.load_checkpoint(CKPT), precision="fp16", device="cuda")
    pts = t(g["pts"]).cuda()
    with torch.no_grad():
        rgb = net.forward_points(pts)                                  # fused encode + ResMLP (tcgen05)
        rgb_api = net(E.PositionalEmbedder(L=10)(pts))                # the reference's call sequence
        net32 = E.compat.models_from_checkpoint(E.compat.load_checkpoint(CKPT), precision="fp32", device="cuda")[0]
        rgb32 = net32(E.PositionalEmbedder(L=10)(pts))
    ref = t(g["rgb"])
    assert float((rgb32.cpu() - ref).abs().max()) < 2e-5              # fp32 path: reference-exact
    assert float((rgb.cpu() - ref).abs().max()) <= 2e-3
    assert float((rgb_api.cpu() - ref).abs().max()) <= 2e-3
    sd = {k[len("nerf_sd::"):]: t(g[k]) for k in g.files if k.startswith("nerf_sd::")}
    nerf, _ = E.compat.models_from_checkpoint({'network_fn_state_dict': sd},
                                              model=E.NeRF(8, 256, 63, 27, 5, [4], True, precision="fp16"),
                                              device="cuda")
    with torch.no_grad():
        out = nerf(t(g["nerf_x"]).cuda())
    assert float((out.cpu() - t(g["nerf_out"])).abs().max()) <= 2e-3


_TORCH2_HOOK_ATTRS = ("_forward_hooks_with_kwargs", "_forward_hooks_always_called", "_forward_pre_hooks_with_kwargs",
                      "_backward_pre_hooks", "_state_dict_pre_hooks", "_load_state_dict_post_hooks",
                      "_load_state_dict_pre_hooks", "_non_persistent_buffers_set", "_is_full_backward_hook")


def test_torch1_pickles_load(E, O):
    """the released R2L / NeRF checkpoints were pickled by torch 1.x, whose modules lack the hook
    dicts torch 2.x expects; nn.Module.__setstate__ back-fills them, so the overrides must call it.  Simulated by
    stripping those attributes from every module's state before pickling."""
    import copy
    import pickle
    for net in (E.NeRF_v3_2(O.r2l_args(netdepth=6), 252, 3), E.NeRF(8, 256, 63, 27, 5, [4], True)):
        old = copy.deepcopy(net)
        for m in old.modules():
            for a in _TORCH2_HOOK_ATTRS:
                m.__dict__.pop(a, None)
        for a in ("precision", "_packed"):          # a module pickled by the reference class has neither
            old.__dict__.pop(a, None)
        new = pickle.loads(pickle.dumps(old))
        assert new.precision in ("fp16", "bf16", "fp32") and new._packed == {}
        new.load_state_dict(net.state_dict())       # AttributeError before the fix
        assert all(torch.equal(v, net.state_dict()[k]) for k, v in new.state_dict().items())
        new.register_forward_hook(lambda *a: None)
