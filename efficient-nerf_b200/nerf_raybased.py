"""B200-native stand-in for the hot-path classes of the reference's model/nerf_raybased.py.

Same class names, constructor arguments, attribute names and state_dict keys as the reference:
  NeRF              model/nerf_raybased.py:337-401
  ResMLP            model/nerf_raybased.py:443-465
  NeRF_v3_2         model/nerf_raybased.py:480-544
  PointSampler      model/nerf_raybased.py:76-126
  PositionalEmbedder model/nerf_raybased.py:191-208
  raw2outputs       model/nerf_raybased.py:226-295
so checkpoints load unchanged; `forward` runs the fused tcgen05 kernels (precision 'fp16' or 'bf16'
operands, fp32 accumulate) or the fp32 CUDA-core path (precision 'fp32', any architecture).
Inference only: inputs that require grad raise (training stays on the reference path).
"""
import ctypes

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import _lib
from .run_nerf_raybased_helpers import Embedder, get_embedder, _embed  # noqa: F401 (re-exported)
from .run_nerf_raybased_helpers import raw2outputs as _raw2outputs


def raw2outputs(raw, z_vals, rays_d, raw_noise_std=0, white_bkgd=False, pytest=False, global_step=-1, print=print):
    """The module's own twin of raw2outputs (model/nerf_raybased.py:226-295: same math, `global_step` / `print` in
    place of `verbose`)."""
    return _raw2outputs(raw, z_vals, rays_d, raw_noise_std, white_bkgd, pytest)


def batchify(fn, chunk):
    """model/nerf_raybased.py:298-309."""
    from .render import batchify as _b
    return _b(fn, chunk)


def run_network(inputs, viewdirs, fn, embed_fn, embeddirs_fn, netchunk=1024 * 64):
    """model/nerf_raybased.py:312-334."""
    from .render import run_network as _r
    return _r(inputs, viewdirs, fn, embed_fn, embeddirs_fn, netchunk)

DEFAULT_PRECISION = "fp16"   # 'fp16' | 'bf16' | 'fp32'
_DTYPE_CODE = {"fp16": 0, "bf16": 1}


def _check_precision(p):
    if p not in ("fp16", "bf16", "fp32"):
        raise ValueError(f"precision must be 'fp16', 'bf16' or 'fp32', got {p!r}")
    return p


def _params_key(module):
    return tuple((p.data_ptr(), p._version, tuple(p.shape)) for p in module.parameters())


def _check_infer_input(x, name="input"):
    if not isinstance(x, torch.Tensor):
        raise TypeError(f"{name} must be a torch.Tensor")
    if not x.is_cuda:
        raise RuntimeError(f"{name} is on {x.device}: the r2l_b200 modules run on CUDA only (no CPU fallback)")
    if x.requires_grad and torch.is_grad_enabled():
        raise RuntimeError("r2l_b200 modules are inference-only; call under torch.no_grad()")


def _rows(t, dev):
    """2-D fp32 CUDA view with unit inner stride (row stride is passed to the kernel); else a copy."""
    if (isinstance(t, torch.Tensor) and t.is_cuda and t.device == dev and t.dtype == torch.float32 and t.dim() == 2
            and t.stride(1) == 1 and not (t.requires_grad and torch.is_grad_enabled())):
        return t.detach()
    return _lib.as_f32_cuda(t, dev)


class _Handle:
    """Owns one packed-weight handle of the C library."""

    def __init__(self, h):
        self.h = h

    def status(self):
        """r2l_mlp_status: 0 healthy, 4 a kernel watchdog fired, 5 a finished launch produced non-finite outputs (an
        activation left the fp16 range).  Reads the handle's mapped record: synchronise the stream first."""
        rec = (ctypes.c_uint * 8)()
        return int(_lib.load().r2l_mlp_status(self.h, rec)), list(rec)

    def __del__(self):
        try:
            if self.h:
                _lib.load().r2l_mlp_destroy(self.h)
        except Exception:
            pass
        self.h = None


# ----------------------------------------------------------------------------- fp32 path
def _pad4(x):
    """[M, K] -> contiguous [M, K4] with zero padded columns (K4 % 4 == 0, 16-byte aligned rows)."""
    K = x.shape[-1]
    K4 = (K + 3) // 4 * 4
    if K4 == K and x.is_contiguous() and x.data_ptr() % 16 == 0:
        return x
    out = torch.zeros((x.shape[0], K4), dtype=torch.float32, device=x.device)
    out[:, :K] = x
    return out


def _linear_fp32(x, weight, bias, act=0, residual=None, scale=1.0):
    """act(x @ (scale*W)^T + scale*b [+ residual]) through r2l_linear_fp32.  x [M, K]."""
    M, K = x.shape
    N = weight.shape[0]
    x4 = _pad4(x)
    w4 = _pad4(weight.detach().to(torch.float32) * scale if scale != 1.0 else weight.detach().to(torch.float32))
    b = None
    if bias is not None:
        b = (bias.detach().to(torch.float32) * scale if scale != 1.0 else bias.detach().to(torch.float32)).contiguous()
    y = torch.empty((M, N), dtype=torch.float32, device=x.device)
    r = residual.contiguous() if residual is not None else None
    max_rows = 4 * 1024 * 1024
    with torch.cuda.device(x.device):
        for m0 in range(0, M, max_rows):
            m1 = min(M, m0 + max_rows)
            _lib.call("r2l_linear_fp32", m1 - m0, N, x4.shape[1], _lib.ptr(x4[m0:m1]), x4.stride(0), _lib.ptr(w4),
                      w4.stride(0), _lib.ptr(b), _lib.ptr(y[m0:m1]), y.stride(0), act,
                      _lib.ptr(r[m0:m1]) if r is not None else _lib.ptr(None), r.stride(0) if r is not None else 0,
                      _lib.stream_ptr(x.device))
    return y


def _act_code(m):
    if m is None:
        return 0, None
    if isinstance(m, nn.ReLU):
        return 1, None
    if isinstance(m, nn.Sigmoid):
        return 2, None
    return 0, m  # applied with torch afterwards (still on the GPU)


def _run_sequential_fp32(seq, x):
    """Execute an nn.Sequential of Linear / activation / ResMLP modules with the fp32 kernel."""
    mods = list(seq.children()) if isinstance(seq, nn.Sequential) else [seq]
    i = 0
    while i < len(mods):
        m = mods[i]
        if isinstance(m, nn.Linear):
            nxt = mods[i + 1] if i + 1 < len(mods) else None
            code, post = (0, None)
            if nxt is not None and not isinstance(nxt, (nn.Linear, ResMLP)):
                code, post = _act_code(nxt)
                i += 1
            x = _linear_fp32(x, m.weight, m.bias, code)
            if post is not None:
                x = post(x)
        elif isinstance(m, ResMLP):
            x = m._forward_fp32(x)
        elif isinstance(m, nn.Identity):
            pass
        else:
            x = m(x)
        i += 1
    return x


# ----------------------------------------------------------------------------- NeRF
class NeRF(nn.Module):

    def __init__(self, D=8, W=256, input_ch=3, input_ch_views=3, output_ch=4, skips=[4], use_viewdirs=False,
                 precision=None):
        super(NeRF, self).__init__()
        self.D = D
        self.W = W
        self.input_ch = input_ch
        self.input_ch_views = input_ch_views
        self.skips = skips
        self.use_viewdirs = use_viewdirs
        # construction order = the reference's, so a seeded init draws identical weights (model:357-375)
        self.pts_linears = nn.ModuleList([nn.Linear(input_ch, W)] + [
            nn.Linear(W, W) if i not in self.skips else nn.Linear(W + input_ch, W) for i in range(D - 1)
        ])
        self.views_linears = nn.ModuleList([nn.Linear(input_ch_views + W, W // 2)])
        if use_viewdirs:
            self.feature_linear = nn.Linear(W, W)
            self.alpha_linear = nn.Linear(W, 1)
            self.rgb_linear = nn.Linear(W // 2, 3)
        else:
            self.output_linear = nn.Linear(W, output_ch)
        self.precision = _check_precision(precision or DEFAULT_PRECISION)
        self._packed = {}


    def __getstate__(self):
        d = self.__dict__.copy()
        d['_packed'] = {}   # C handles are process-local; they are rebuilt on demand after unpickling
        return d

    def __setstate__(self, state):
        # also reached when a module pickled by the REFERENCE class is loaded through compat's aliases; the base
        # method back-fills the hook dicts that torch-1.x pickles (the released checkpoints) do not carry
        super(NeRF, self).__setstate__(state)
        self.__dict__.setdefault('precision', DEFAULT_PRECISION)
        self.__dict__['_packed'] = {}

    def _replicate_for_data_parallel(self):
        # nn.DataParallel shallow-copies __dict__: without this every replica would share (and evict, from concurrent
        # threads) one dict of packed-weight handles keyed by ITS OWN parameter copies
        replica = super()._replicate_for_data_parallel()
        replica._packed = {}
        return replica

    # -- tensor-core path -------------------------------------------------------------------
    def supports_tensor_core_path(self):
        return (self.D == 8 and self.W == 256 and self.input_ch == 63 and self.input_ch_views == 27
                and list(self.skips) == [4] and self.use_viewdirs)

    def packed_handle(self, precision=None):
        """Packed 16-bit weights for the fused kernel (re-packed when parameters change)."""
        precision = _check_precision(precision or self.precision)
        if precision == "fp32" or not self.supports_tensor_core_path():
            raise RuntimeError("this NeRF configuration/precision has no tensor-core handle")
        key = (precision, _params_key(self))
        ent = self._packed.get(precision)
        if ent is not None and ent[0] == key:
            return ent[1]
        dev = self.alpha_linear.weight.device
        if dev.type != "cuda":
            raise RuntimeError("NeRF parameters must be on a CUDA device (no CPU fallback)")
        f = lambda t: t.detach().to(torch.float32).contiguous()
        pw = [f(l.weight) for l in self.pts_linears]
        pb = [f(l.bias) for l in self.pts_linears]
        others = [f(self.views_linears[0].weight), f(self.views_linears[0].bias), f(self.feature_linear.weight),
                  f(self.feature_linear.bias), f(self.alpha_linear.weight), f(self.alpha_linear.bias),
                  f(self.rgb_linear.weight), f(self.rgb_linear.bias)]
        out = ctypes.c_void_p(0)
        with torch.cuda.device(dev):
            _lib.call("r2l_nerf_create", ctypes.byref(out), _DTYPE_CODE[precision], _lib.ptr_array(pw),
                      _lib.ptr_array(pb), *[_lib.ptr(t) for t in others], _lib.stream_ptr(dev))
        h = _Handle(out)
        self._packed[precision] = (key, h)
        self._apply_far(h)
        return h

    # -- far-sample sigma fix-up (csrc/nerf_far.cu) ------------------------------------------
    def set_far_fixup(self, enabled=True, abs_band=None, rel_band=None):
        """raw2outputs gives a ray's LAST sample a 1e10 interval (main.py:579-581), so its alpha is a step function of
        sign(sigma).  With the fix-up on (default) forward_samples re-evaluates, in fp32, the far samples whose 16-bit
        sigma lies inside the guard band |sigma| < max(abs_band, rel_band * sum |w_a| relu(h7)) — bit-identical to the
        precision='fp32' path there.  `enabled=False` returns the tensor-core kernels' own sigma everywhere."""
        self.__dict__['_far'] = (bool(enabled), abs_band, rel_band)
        for _, h in self._packed.values():
            self._apply_far(h)

    def _apply_far(self, h):
        en, a, r = self.__dict__.get('_far', (True, None, None))
        _lib.call("r2l_nerf_far_fixup", h.h, 1 if en else 0, -1.0 if a is None else float(a), -1.0 if r is None else float(r))

    def range_status(self):
        """(code, record) of the packed handle after synchronising: code 5 = the last launches overflowed the 16-bit
        operand range (outputs non-finite); the next forward raises the same error once."""
        torch.cuda.synchronize(self.alpha_linear.weight.device)
        return self.packed_handle().status()

    def far_flagged(self):
        """Rays the last forward_samples flagged for the fp32 far-sample fix-up (synchronises the stream)."""
        n = ctypes.c_longlong(0)
        h = self.packed_handle()
        dev = self.alpha_linear.weight.device
        with torch.cuda.device(dev):
            _lib.call("r2l_nerf_far_count", h.h, ctypes.byref(n), _lib.stream_ptr(dev))
        return int(n.value)

    def forward_samples(self, rays_o, rays_d, viewdirs, z_vals):
        """Fused encode + MLP: raw [N, S, 4] for points o + d*z (main.py:65-87 + model:377-401)."""
        _check_infer_input(z_vals, "z_vals")
        dev = z_vals.device
        N, S = z_vals.shape
        h = self.packed_handle()
        ro, rd, vd = (_rows(t, dev) for t in (rays_o, rays_d, viewdirs))
        z = _lib.as_f32_cuda(z_vals, dev)
        raw = torch.empty((N, S, 4), dtype=torch.float32, device=dev)
        evs = self.__dict__.get('_mlp_events')      # bench.py: CUDA events around the MLP launches of a step
        with torch.cuda.device(dev):
            if evs is not None:
                e0 = torch.cuda.Event(enable_timing=True)
                e0.record()
            _lib.call("r2l_nerf_forward", h.h, N, S, _lib.ptr(ro), ro.stride(0), _lib.ptr(rd), rd.stride(0),
                      _lib.ptr(vd), vd.stride(0), _lib.ptr(z), _lib.ptr(raw), _lib.stream_ptr(dev))
            if evs is not None:
                e1 = torch.cuda.Event(enable_timing=True)
                e1.record()
                evs.append((e0, e1))
        return raw

    def render_samples(self, rays_o, rays_d, viewdirs, z_vals, white_bkgd=False, want_weights=True):
        """Fused encode + MLP + raw2outputs (main.py:707-709 / 738-741 with raw_noise_std = 0): returns rgb_map,
        disp_map, acc_map, weights (or None), depth_map — one library call (r2l_nerf_render); the same bits as
        raw2outputs(forward_samples(...)).  With set_fused_compositing(True) (or R2L_NERF_FUSED=1) and 64 / 128 / 192 /
        256 samples the CTA-pair kernel composites the rays itself and raw [N, S, 4] is never written: 1.3 GB less HBM
        traffic per 400x400 frame, bit-identical, ~1 % slower — hence not the default."""
        _check_infer_input(z_vals, "z_vals")
        dev = z_vals.device
        N, S = z_vals.shape
        h = self.packed_handle()
        ro, rd, vd = (_rows(t, dev) for t in (rays_o, rays_d, viewdirs))
        z = _lib.as_f32_cuda(z_vals, dev)
        rgb_map = torch.empty((N, 3), dtype=torch.float32, device=dev)
        disp_map = torch.empty((N,), dtype=torch.float32, device=dev)
        acc_map = torch.empty((N,), dtype=torch.float32, device=dev)
        depth_map = torch.empty((N,), dtype=torch.float32, device=dev)
        weights = torch.empty((N, S), dtype=torch.float32, device=dev) if want_weights else None
        evs = self.__dict__.get('_mlp_events')      # bench.py: CUDA events around the MLP launches of a step
        with torch.cuda.device(dev):
            if evs is not None:
                e0 = torch.cuda.Event(enable_timing=True)
                e0.record()
            _lib.call("r2l_nerf_render", h.h, N, S, _lib.ptr(ro), ro.stride(0), _lib.ptr(rd), rd.stride(0),
                      _lib.ptr(vd), vd.stride(0), _lib.ptr(z), int(bool(white_bkgd)), _lib.ptr(rgb_map),
                      _lib.ptr(disp_map), _lib.ptr(acc_map), _lib.ptr(weights), _lib.ptr(depth_map),
                      _lib.stream_ptr(dev))
            if evs is not None:
                e1 = torch.cuda.Event(enable_timing=True)
                e1.record()
                evs.append((e0, e1))
        return rgb_map, disp_map, acc_map, weights, depth_map

    def set_fused_compositing(self, enabled=True):
        """True: render_samples composites inside the MLP kernel where it can; False (default): MLP -> raw workspace ->
        raw2outputs inside the same library call."""
        _lib.call("r2l_nerf_render_mode", self.packed_handle().h, 1 if enabled else 0)
        self.__dict__['_fused_comp'] = bool(enabled)

    # -- nn.Module API ---------------------------------------------------------------------
    def forward(self, x):
        _check_infer_input(x, "x")
        lead = x.shape[:-1]
        x2 = x.detach().to(torch.float32).reshape(-1, x.shape[-1])
        if self.precision != "fp32" and self.supports_tensor_core_path() and x2.shape[-1] >= 90:
            x2 = x2.contiguous()
            M = x2.shape[0]
            h = self.packed_handle()
            out = torch.empty((M, 4), dtype=torch.float32, device=x2.device)
            with torch.cuda.device(x2.device):
                _lib.call("r2l_nerf_forward_embedded", h.h, M, _lib.ptr(x2), x2.stride(0), _lib.ptr(out),
                          _lib.stream_ptr(x2.device))
            return out.reshape(*lead, 4)
        return self._forward_fp32(x2).reshape(*lead, -1)

    def _forward_fp32(self, x):
        input_pts, input_views = torch.split(x, [self.input_ch, self.input_ch_views], dim=-1)
        h = input_pts
        for i, l in enumerate(self.pts_linears):
            h = _linear_fp32(h, l.weight, l.bias, 1)
            if i in self.skips:
                h = torch.cat([input_pts, h], -1)
        if self.use_viewdirs:
            alpha = _linear_fp32(h, self.alpha_linear.weight, self.alpha_linear.bias, 0)
            feature = _linear_fp32(h, self.feature_linear.weight, self.feature_linear.bias, 0)
            h = torch.cat([feature, input_views], -1)
            for l in self.views_linears:
                h = _linear_fp32(h, l.weight, l.bias, 1)
            rgb = _linear_fp32(h, self.rgb_linear.weight, self.rgb_linear.bias, 0)
            return torch.cat([rgb, alpha], -1)
        return _linear_fp32(h, self.output_linear.weight, self.output_linear.bias, 0)


# ----------------------------------------------------------------------------- R2L
class ResMLP(nn.Module):

    def __init__(self, width, inact=nn.ReLU(True), outact=None, res_scale=1, n_learnable=2):
        '''inact is the activation func within block. outact is the activation func right before output'''
        super(ResMLP, self).__init__()
        m = [nn.Linear(width, width)]
        for _ in range(n_learnable - 1):
            if inact is not None:
                m += [inact]
            m += [nn.Linear(width, width)]
        self.body = nn.Sequential(*m)
        self.res_scale = res_scale
        self.outact = outact

    def _forward_fp32(self, x):
        mods = list(self.body.children())
        linears = [i for i, m in enumerate(mods) if isinstance(m, nn.Linear)]
        y = x
        for j, li in enumerate(linears):
            lin = mods[li]
            last = j == len(linears) - 1
            nxt = mods[li + 1] if li + 1 < len(mods) and not isinstance(mods[li + 1], nn.Linear) else None
            code, post = _act_code(nxt)
            if last:
                # x = body(x) * res_scale + x   (model:462)
                y = _linear_fp32(y, lin.weight, lin.bias, 0, residual=x, scale=float(self.res_scale))
            else:
                y = _linear_fp32(y, lin.weight, lin.bias, code)
                if post is not None:
                    y = post(y)
        if self.outact is not None:
            y = self.outact(y)
        return y

    def forward(self, x):
        _check_infer_input(x, "x")
        lead = x.shape[:-1]
        return self._forward_fp32(x.detach().to(torch.float32).reshape(-1, x.shape[-1])).reshape(*lead, -1)


def get_activation(act):
    if act.lower() == 'relu':
        func = nn.ReLU(inplace=True)
    elif act.lower() == 'lrelu':
        func = nn.LeakyReLU(inplace=True)
    elif act.lower() == 'none':
        func = None
    else:
        raise NotImplementedError
    return func


class LazyPoints:
    """Deferred PointSampler.sample_test output (PointSampler(..., lazy=True)): remembers sampler and pose(s) so that
    NeRF_v3_2 can generate the rays INSIDE the fused kernel (r2l_resmlp_render: no [H*W, n_sample*3] tensor at all);
    anything else that touches it gets the materialised points."""

    def __init__(self, sampler, c2ws):
        self.sampler, self.c2ws = sampler, c2ws      # c2ws: [P, 3, 4] on the sampler's device
        self._mat = None

    @property
    def shape(self):
        return torch.Size((self.c2ws.shape[0] * self.sampler.H * self.sampler.W, self.sampler.n_sample * 3))

    def dim(self):
        return 2

    @property
    def device(self):
        return self.c2ws.device

    @property
    def dtype(self):
        return torch.float32

    @property
    def requires_grad(self):
        return False

    def materialize(self):
        if self._mat is None:
            self._mat = self.sampler.sample_test_batch(self.c2ws, lazy=False)
        return self._mat

    def __getattr__(self, name):
        return getattr(self.materialize(), name)

    def __getitem__(self, idx):
        return self.materialize()[idx]

    @classmethod
    def __torch_function__(cls, func, types, args=(), kwargs=None):
        conv = lambda a: a.materialize() if isinstance(a, LazyPoints) else a
        return func(*[conv(a) for a in args], **{k: conv(v) for k, v in (kwargs or {}).items()})


class LazyEmbedding:
    """Deferred PositionalEmbedder output: remembers the sampled points so that NeRF_v3_2 can run the
    fused encode+MLP kernel; anything else that touches it gets the materialised [N, dim*(2L+1)] tensor."""

    def __init__(self, pts, L, include_input):
        self.pts, self.L, self.include_input = pts, L, include_input
        self._mat = None

    @property
    def shape(self):
        d = self.pts.shape[-1] * (2 * self.L + (1 if self.include_input else 0))
        return torch.Size(tuple(self.pts.shape[:-1]) + (d,))

    @property
    def device(self):
        return self.pts.device

    @property
    def dtype(self):
        return torch.float32

    def materialize(self):
        if self._mat is None:
            pts = self.pts.materialize() if isinstance(self.pts, LazyPoints) else self.pts
            self._mat = _embed(pts, self.L, self.include_input, 1)
        return self._mat

    def __getattr__(self, name):
        return getattr(self.materialize(), name)

    def __getitem__(self, idx):
        return self.materialize()[idx]

    @classmethod
    def __torch_function__(cls, func, types, args=(), kwargs=None):
        conv = lambda a: a.materialize() if isinstance(a, LazyEmbedding) else a
        return func(*[conv(a) for a in args], **{k: conv(v) for k, v in (kwargs or {}).items()})


class NeRF_v3_2(nn.Module):
    '''Based on NeRF_v3, move positional embedding out'''

    def __init__(self, args, input_dim, output_dim, precision=None):
        super(NeRF_v3_2, self).__init__()
        self.args = args
        D, W = args.netdepth, args.netwidth
        if getattr(args, 'layerwise_netwidths', ''):
            Ws = [int(x) for x in args.layerwise_netwidths.split(',')] + [3]
        else:
            Ws = [W] * (D - 1) + [3]
        act = get_activation(args.act)
        self.input_dim = input_dim
        self.head = nn.Sequential(*[nn.Linear(input_dim, Ws[0]), act])
        # the reference first builds (and then discards) a plain MLP body; doing the same keeps the
        # RNG stream, hence a seeded random init, identical (model:503-505)
        body = []
        for i in range(1, D - 1):
            body += [nn.Linear(Ws[i - 1], Ws[i]), act]
        if hasattr(args, 'trial'):
            inact = get_activation(args.trial.inact)
            outact = get_activation(args.trial.outact)
            if args.trial.body_arch in ['resmlp']:
                n_block = (D - 2) // 2
                if args.trial.n_block > 0:
                    n_block = args.trial.n_block
                body = [
                    ResMLP(W, inact=inact, outact=outact, res_scale=args.trial.res_scale,
                           n_learnable=args.trial.n_learnable) for _ in range(n_block)
                ]
            elif args.trial.body_arch in ['mlp']:
                body = []
                for i in range(1, D - 1):
                    body += [nn.Linear(Ws[i - 1], Ws[i]), act]
        self.body = nn.Sequential(*body)
        self.tail = nn.Linear(input_dim, output_dim) if args.linear_tail else nn.Sequential(
            *[nn.Linear(Ws[D - 2], output_dim), nn.Sigmoid()])
        self.precision = _check_precision(precision or DEFAULT_PRECISION)
        self._packed = {}


    def __getstate__(self):
        d = self.__dict__.copy()
        d['_packed'] = {}   # C handles are process-local; they are rebuilt on demand after unpickling
        return d

    def __setstate__(self, state):
        # also reached when a module pickled by the REFERENCE class (main.py:1534-1536) is loaded through compat; the
        # base method back-fills the hook dicts that torch-1.x pickles (the released R2L checkpoints) do not carry
        super(NeRF_v3_2, self).__setstate__(state)
        self.__dict__.setdefault('precision', DEFAULT_PRECISION)
        self.__dict__['_packed'] = {}
        if 'input_dim' not in self.__dict__:
            self.__dict__['input_dim'] = self.head[0].in_features

    def _replicate_for_data_parallel(self):
        # nn.DataParallel shallow-copies __dict__: without this every replica would share (and evict, from concurrent
        # threads) one dict of packed-weight handles keyed by ITS OWN parameter copies
        replica = super()._replicate_for_data_parallel()
        replica._packed = {}
        return replica

    # -- tensor-core path -------------------------------------------------------------------
    def _tc_config(self):
        """(n_points, blocks) if the fused kernel covers this architecture, else None."""
        a = self.args
        if not hasattr(a, 'trial') or a.trial.body_arch not in ['resmlp'] or a.linear_tail:
            return None
        if a.netwidth != 256 or getattr(a, 'layerwise_netwidths', '') or a.act.lower() != 'relu':
            return None
        if a.trial.inact.lower() != 'relu' or a.trial.outact.lower() != 'none' or a.trial.n_learnable != 2:
            return None
        if self.input_dim % 63 != 0 or (self.input_dim // 63) % 4 != 0 or self.input_dim // 63 > 256:
            return None
        blocks = [m for m in self.body.children()]
        if not blocks or not all(isinstance(b, ResMLP) for b in blocks):
            return None
        tail0 = self.tail[0]
        if tail0.out_features != 3 or tail0.in_features != 256:
            return None
        return self.input_dim // 63, blocks

    def supports_tensor_core_path(self):
        return self._tc_config() is not None

    def packed_handle(self, precision=None):
        precision = _check_precision(precision or self.precision)
        cfg = self._tc_config()
        if precision == "fp32" or cfg is None:
            raise RuntimeError("this NeRF_v3_2 configuration/precision has no tensor-core handle")
        n_points, blocks = cfg
        key = (precision, _params_key(self))
        ent = self._packed.get(precision)
        if ent is not None and ent[0] == key:
            return ent[1]
        dev = self.head[0].weight.device
        if dev.type != "cuda":
            raise RuntimeError("NeRF_v3_2 parameters must be on a CUDA device (no CPU fallback)")
        f = lambda t: t.detach().to(torch.float32).contiguous()
        w1 = [f(b.body[0].weight) for b in blocks]
        b1 = [f(b.body[0].bias) for b in blocks]
        w2 = [f(b.body[2].weight) for b in blocks]
        b2 = [f(b.body[2].bias) for b in blocks]
        res_scales = {float(b.res_scale) for b in blocks}
        if len(res_scales) != 1:
            raise RuntimeError("blocks with different res_scale are not supported by the fused kernel")
        hw, hb = f(self.head[0].weight), f(self.head[0].bias)
        tw, tb = f(self.tail[0].weight), f(self.tail[0].bias)
        out = ctypes.c_void_p(0)
        with torch.cuda.device(dev):
            _lib.call("r2l_resmlp_create", ctypes.byref(out), _DTYPE_CODE[precision], n_points, len(blocks),
                      _lib.ptr(hw), _lib.ptr(hb), _lib.ptr_array(w1), _lib.ptr_array(b1), _lib.ptr_array(w2),
                      _lib.ptr_array(b2), res_scales.pop(), _lib.ptr(tw), _lib.ptr(tb), 1,
                      int(bool(self.args.use_residual)), _lib.stream_ptr(dev))
        h = _Handle(out)
        self._packed[precision] = (key, h)
        return h

    def range_status(self):
        """(code, record) of the packed handle after synchronising: code 5 = the last launches overflowed the 16-bit
        operand range (outputs non-finite); the next forward raises the same error once."""
        torch.cuda.synchronize(self.head[0].weight.device)
        return self.packed_handle().status()

    def forward_points(self, pts):
        """Fused PositionalEmbedder(L=10) + network: pts [N, n_points*3] -> rgb [N, 3]."""
        _check_infer_input(pts, "pts")
        h = self.packed_handle()
        p = _lib.as_f32_cuda(pts).reshape(-1, pts.shape[-1])
        N = p.shape[0]
        rgb = torch.empty((N, 3), dtype=torch.float32, device=p.device)
        with torch.cuda.device(p.device):
            _lib.call("r2l_resmlp_forward", h.h, N, _lib.ptr(p), p.stride(0), _lib.ptr(rgb), _lib.stream_ptr(p.device))
        return rgb

    def render_poses(self, sampler, c2ws, rows=None, frame=None):
        """PointSampler.sample_test + PositionalEmbedder(L=10) + network in ONE kernel (main.py:297-309): the head
        generates its rays from pixel index and pose, no points tensor.  c2ws [3|4, 4] or [P, 3|4, 4]; `rows` =
        (start, stop) of the pose-major ray range [P*H*W] to render (default: all).  With `frame`
        (sharding.PeerFrame, one pose) this rank's rows are stored into every GPU's frame buffer (call
        frame.publish() afterwards).  Bit-identical to forward_points(sampler.sample_test(c2w))."""
        if sampler.n_sample * 63 != self.input_dim:
            raise ValueError(f"the sampler draws {sampler.n_sample} points per ray, the model takes {self.input_dim // 63}")
        h = self.packed_handle()
        dev = sampler.z_vals.device
        c = _lib.as_f32_cuda(c2ws, dev, "c2w")
        c = c.reshape(-1, c.shape[-2], 4)[:, :3, :4].contiguous()
        P, n_all = c.shape[0], c.shape[0] * sampler.H * sampler.W
        if frame is not None:
            if P != 1 or frame.n_rays != n_all:
                raise ValueError("a PeerFrame holds one frame of this sampler's size")
            rows = (frame.row0, frame.row1)
        r0, r1 = (0, n_all) if rows is None else (int(rows[0]), int(rows[1]))
        if not (0 <= r0 <= r1 <= n_all):
            raise ValueError(f"rows {rows} outside [0, {n_all}]")
        z = sampler.z_vals.contiguous()
        rgb = None if frame is not None else torch.empty((r1 - r0, 3), dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            _lib.call("r2l_resmlp_render", h.h, P, sampler.H, sampler.W, sampler.focal, _lib.ptr(c), _lib.ptr(z),
                      sampler.n_sample, r0, r1 - r0, _lib.ptr(rgb), frame.ptrs if frame is not None else None,
                      frame.world_size if frame is not None else 0, r0 if frame is not None else 0,
                      _lib.stream_ptr(dev))
        return frame.buf if frame is not None else rgb

    def forward_points_gather(self, pts, frame):
        """forward_points for ONE RANK'S ray block of a sharded frame, with the tile gather fused into the kernel:
        rgb rows land in `frame` (sharding.PeerFrame: the same [n_rays, 3] symmetric-memory buffer on every GPU of the
        box) at this rank's row offset ON EVERY RANK by peer-to-peer stores.  Call frame.publish() afterwards (the
        cross-GPU barrier that replaces the NCCL all-gather); returns the local view of the whole frame."""
        _check_infer_input(pts, "pts")
        h = self.packed_handle()
        p = _lib.as_f32_cuda(pts).reshape(-1, pts.shape[-1])
        N = p.shape[0]
        if N != frame.row1 - frame.row0:
            raise ValueError(f"this rank owns rows [{frame.row0}, {frame.row1}) of the frame, got {N} rays")
        with torch.cuda.device(p.device):
            _lib.call("r2l_resmlp_forward_gather", h.h, N, _lib.ptr(p), p.stride(0), frame.ptrs, frame.world_size,
                      frame.row0, _lib.stream_ptr(p.device))
        return frame.buf

    # -- nn.Module API ---------------------------------------------------------------------
    def forward(self, x):  # x: embedded position coordinates
        tc = self.precision != "fp32" and self.supports_tensor_core_path()
        if isinstance(x, LazyEmbedding):
            if tc and x.L == 10 and x.include_input and x.pts.shape[-1] * 21 == self.input_dim:
                if isinstance(x.pts, LazyPoints):      # rays generated inside the kernel
                    return self.render_poses(x.pts.sampler, x.pts.c2ws)
                return self.forward_points(x.pts)
            x = x.materialize()
        _check_infer_input(x, "x")
        if x.shape[-1] != self.input_dim:  # [N, C, H, W]
            x = x.permute(0, 2, 3, 1)
        lead = x.shape[:-1]
        x2 = x.detach().to(torch.float32).reshape(-1, x.shape[-1])
        if tc:
            x2 = x2.contiguous()
            N = x2.shape[0]
            h = self.packed_handle()
            rgb = torch.empty((N, 3), dtype=torch.float32, device=x2.device)
            with torch.cuda.device(x2.device):
                _lib.call("r2l_resmlp_forward_embedded", h.h, N, _lib.ptr(x2), x2.stride(0), _lib.ptr(rgb),
                          _lib.stream_ptr(x2.device))
            return rgb.reshape(*lead, 3)
        return self._forward_fp32(x2).reshape(*lead, -1)

    def _forward_fp32(self, x):
        x = _run_sequential_fp32(self.head, x)
        y = _run_sequential_fp32(self.body, x)
        x = y + x if self.args.use_residual else y
        return _run_sequential_fp32(self.tail, x)


# ----------------------------------------------------------------------------- R2L input pipeline
class PointSampler():

    def __init__(self, H, W, focal, n_sample, near, far, lazy=False):
        _lib.require_cuda()
        self.H, self.W, self.focal = int(H), int(W), float(focal)
        self.n_sample = int(n_sample)
        self.lazy = lazy        # sample_test returns a LazyPoints handle (rays generated inside the fused kernel)
        dev = torch.device("cuda", torch.cuda.current_device())
        t_vals = torch.linspace(0., 1., steps=n_sample).to(dev)  # host linspace, like the reference (model:88-89)
        self.z_vals = near * (1 - t_vals) + far * (t_vals)       # [n_sample]
        self.z_vals_test = self.z_vals[None, :].expand(self.H * self.W, n_sample)
        self.near, self.far = near, far

    def _sample(self, c2w):
        dev = self.z_vals.device
        c = _lib.as_f32_cuda(c2w, dev, "c2w")[:3, :4].contiguous()
        pts = torch.empty((self.H * self.W, self.n_sample * 3), dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            _lib.call("r2l_point_sample", self.H, self.W, self.focal, _lib.ptr(c), _lib.ptr(self.z_vals),
                      self.n_sample, _lib.ptr(pts), _lib.stream_ptr(dev))
        return pts

    def sample_test(self, c2w):  # c2w: [3, 4]
        if self.lazy:
            c = _lib.as_f32_cuda(c2w, self.z_vals.device, "c2w")
            return LazyPoints(self, c[:3, :4].reshape(1, 3, 4).contiguous())
        return self._sample(c2w)  # [H*W, n_sample*3]

    def sample_test_batch(self, c2ws, lazy=None):
        """sample_test for P poses in one launch: c2ws [P, 3|4, 4] -> [P*H*W, n_sample*3] (pose-major), so that one
        fused-MLP launch renders P frames (full waves of 128-ray tiles instead of a ragged last wave per frame)."""
        dev = self.z_vals.device
        c = _lib.as_f32_cuda(c2ws, dev, "c2ws")
        if c.dim() != 3 or c.shape[-1] != 4 or c.shape[-2] < 3:
            raise ValueError(f"c2ws must be [P, 3, 4] or [P, 4, 4], got {tuple(c.shape)}")
        c = c[:, :3, :4].contiguous()
        if self.lazy if lazy is None else lazy:
            return LazyPoints(self, c)
        P = c.shape[0]
        pts = torch.empty((P * self.H * self.W, self.n_sample * 3), dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            _lib.call("r2l_point_sample_batch", P, self.H, self.W, self.focal, _lib.ptr(c), _lib.ptr(self.z_vals),
                      self.n_sample, _lib.ptr(pts), _lib.stream_ptr(dev))
        return pts

    def sample_test2(self, c2w):
        return self._sample(c2w).view(self.H * self.W, self.n_sample, 3)

    def sample_train(self, rays_o, rays_d, perturb, t_rand=None):
        dev = self.z_vals.device
        ro = _lib.as_f32_cuda(rays_o, dev, "rays_o").reshape(-1, 3)
        rd = _lib.as_f32_cuda(rays_d, dev, "rays_d").reshape(-1, 3)
        n, S = ro.shape[0], self.n_sample
        pts = torch.empty((n, S * 3), dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            if perturb > 0.:
                if t_rand is None:
                    t_rand = torch.rand((n, S))  # CPU generator, then uploaded (model:121)
                tr = _lib.as_f32_cuda(t_rand, dev, "t_rand")
                mids = .5 * (self.z_vals[1:] + self.z_vals[:-1])
                upper = torch.cat([mids, self.z_vals[-1:]], -1)
                lower = torch.cat([self.z_vals[:1], mids], -1)
                z = (lower + (upper - lower) * tr).contiguous()
                _lib.call("r2l_points_from_rays", n, S, _lib.ptr(ro), 3, _lib.ptr(rd), 3, _lib.ptr(z), S,
                          _lib.ptr(pts), _lib.stream_ptr(dev))
            else:
                z = self.z_vals.contiguous()
                _lib.call("r2l_points_from_rays", n, S, _lib.ptr(ro), 3, _lib.ptr(rd), 3, _lib.ptr(z), 0,
                          _lib.ptr(pts), _lib.stream_ptr(dev))
        return pts

    def sample_train_cnnstyle(self, rays_o, rays_d, perturb, t_rand=None):
        """CNN-style patches (model/nerf_raybased.py:149-168; `sample_train2` :128-147 is the same code): rays_o /
        rays_d [n_img, ph, pw, 3] -> pts [n_img, ph, pw, n_sample, 3]; with perturb > 0 ONE stratification offset
        per image (t_rand [n_img], drawn on the CPU generator like the reference, or injected)."""
        dev = self.z_vals.device
        ro = _lib.as_f32_cuda(rays_o, dev, "rays_o")
        rd = _lib.as_f32_cuda(rays_d, dev, "rays_d")
        if ro.dim() != 4 or ro.shape[-1] != 3 or rd.shape != ro.shape:
            raise ValueError(f"rays_o / rays_d must be [n_img, ph, pw, 3], got {tuple(ro.shape)} / {tuple(rd.shape)}")
        S = self.n_sample
        n_img = ro.shape[0]
        n = ro.numel() // 3
        pts = torch.empty(tuple(ro.shape[:3]) + (S, 3), dtype=torch.float32, device=dev)
        ro2, rd2 = ro.reshape(-1, 3).contiguous(), rd.reshape(-1, 3).contiguous()
        with torch.cuda.device(dev):
            if perturb > 0.:
                if t_rand is None:
                    t_rand = torch.rand(n_img)  # CPU generator, then uploaded (model:160)
                tr = _lib.as_f32_cuda(t_rand, dev, "t_rand").reshape(n_img, 1)
                mids = .5 * (self.z_vals[1:] + self.z_vals[:-1])
                upper = torch.cat([mids, self.z_vals[-1:]], -1)
                lower = torch.cat([self.z_vals[:1], mids], -1)
                z_img = lower + (upper - lower) * tr                                  # [n_img, S]
                z = z_img[:, None, :].expand(n_img, n // n_img, S).reshape(n, S).contiguous()
                _lib.call("r2l_points_from_rays", n, S, _lib.ptr(ro2), 3, _lib.ptr(rd2), 3, _lib.ptr(z), S,
                          _lib.ptr(pts), _lib.stream_ptr(dev))
            else:
                z = self.z_vals.contiguous()
                _lib.call("r2l_points_from_rays", n, S, _lib.ptr(ro2), 3, _lib.ptr(rd2), 3, _lib.ptr(z), 0,
                          _lib.ptr(pts), _lib.stream_ptr(dev))
        return pts

    sample_train2 = sample_train_cnnstyle   # "kept for back-compatibility" in the reference: identical body

    def _plucker(self, ro, o_stride, rd):
        n = rd.shape[0]
        out = torch.empty((n, 6), dtype=torch.float32, device=rd.device)
        with torch.cuda.device(rd.device):
            _lib.call("r2l_plucker", n, _lib.ptr(ro), o_stride, _lib.ptr(rd), 3, _lib.ptr(out),
                      _lib.stream_ptr(rd.device))
        return out

    def sample_train_plucker(self, rays_o, rays_d):
        """Pluecker coordinates [rays_d, rays_o x rays_d] as the ray representation (model:170-176)."""
        dev = self.z_vals.device
        ro = _lib.as_f32_cuda(rays_o, dev, "rays_o").reshape(-1, 3).contiguous()
        rd = _lib.as_f32_cuda(rays_d, dev, "rays_d").reshape(-1, 3).contiguous()
        if ro.shape != rd.shape:
            raise ValueError("rays_o and rays_d must have the same shape")
        return self._plucker(ro, 3, rd)

    def sample_test_plucker(self, c2w):  # c2w: [3, 4]
        """model:178-190 (the `--plucker` branch of render_path, main.py:296-298): [H*W, 6]."""
        from .run_nerf_raybased_helpers import get_rays
        dev = self.z_vals.device
        c = _lib.as_f32_cuda(c2w, dev, "c2w")[:3, :4].contiguous()
        _, rd = get_rays(self.H, self.W, self.focal, c)
        return self._plucker(c[:3, 3].contiguous(), 0, rd.reshape(-1, 3))


class PositionalEmbedder():

    def __init__(self, L, include_input=True, lazy=False):
        _lib.require_cuda()
        self.weights = 2**torch.linspace(0, L - 1, steps=L).to(torch.device("cuda", torch.cuda.current_device()))
        self.L = int(L)
        self.include_input = include_input
        self.embed_dim = 2 * L + 1 if include_input else 2 * L
        self.lazy = lazy

    def __call__(self, x):
        if self.lazy and x.dim() == 2:
            return LazyEmbedding(x if isinstance(x, LazyPoints) else _lib.as_f32_cuda(x), self.L, self.include_input)
        if isinstance(x, LazyPoints):
            x = x.materialize()
        return _embed(x, self.L, self.include_input, 1)  # [n_ray, dim_pts*(2L+1)]

    def embed_cnnstyle(self, x):
        """model:218-223 (and `embed` :210-216, the same code): x [..., dim] -> [..., dim, 2L+1]; the same numbers as
        __call__, without flattening the feature axes."""
        x = _lib.as_f32_cuda(x)
        y = _embed(x.reshape(-1, x.shape[-1]), self.L, self.include_input, 1)
        return y.reshape(tuple(x.shape) + (self.embed_dim,))

    embed = embed_cnnstyle
