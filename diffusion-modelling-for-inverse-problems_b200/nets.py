"""Drop-in for the reference's `nets.py` score networks (MLP :17-35, MLP2 :37-57, PosteriorScore :143-157).

The modules keep the reference's parameter layout — `state_dict()` keys `0,3,5,7.{weight,bias}` —
and its numerics: tanh is applied twice after the first Linear, because the reference registers its activation a
second time under the name 'act' and `nn.Sequential.forward` walks every registered module.  That is
reproduced here on purpose so reference checkpoints (`current_model.pt`, `diffusion.pt`) load and behave identically.

`forward` on CUDA tensors without autograd runs the fused kernels of libdmip_sm100.so (`dmip_mlp_forward`); when
autograd is recording it runs `dmip_mlp_forward_stash` and registers `dmip_mlp_backward` as its backward (tcgen05 kernels,
first-order: parameter and input gradients), so `loss.backward()` of user-written losses runs on the library too; nets
outside the shapes those kernels serve fall back to the torch module chain with a RuntimeWarning.
`GaussianFourierProjection` / `TemporalMLP*` of the reference are unused by every model (nets.py:62-63) and omitted.
"""
import ctypes as C
import warnings

import torch
from torch import nn

from . import _lib

_WARNED_EAGER = [False]


class DmipMlpGrad(C.Structure):
    _fields_ = [("net", _lib.DmipMlp), ("n", C.c_int64), ("x_dim", C.c_int32), ("cond_dim", C.c_int32),
                ("x", C.c_void_p), ("cond", C.c_void_p), ("t", C.c_void_p), ("out", C.c_void_p),
                ("grad_out", C.c_void_p), ("grad_params", C.c_void_p), ("grad_in", C.c_void_p),
                ("workspace", C.c_void_p), ("workspace_bytes", C.c_size_t)]


def _bind_grad():
    L = _lib.require_gpu()
    if not getattr(L, "_mlp_grad_bound", False):
        L.dmip_mlp_grad_workspace_bytes.restype = C.c_size_t
        L.dmip_mlp_grad_workspace_bytes.argtypes = [C.POINTER(DmipMlpGrad)]
        L.dmip_mlp_forward_stash.restype = C.c_int
        L.dmip_mlp_forward_stash.argtypes = [C.POINTER(DmipMlpGrad), C.c_void_p]
        L.dmip_mlp_backward.restype = C.c_int
        L.dmip_mlp_backward.argtypes = [C.POINTER(DmipMlpGrad), C.c_void_p]
        L.dmip_loss_grad_floats.restype = C.c_size_t
        L.dmip_loss_grad_floats.argtypes = [C.POINTER(_lib.DmipMlp)]
        L._mlp_grad_bound = True
    return L


def _grad_desc(net, x, cond, t, keep):
    d = DmipMlpGrad()
    d.net = _lib.mlp_desc(net, keep)
    d.n = x.shape[0]
    d.x_dim = x.shape[1]
    d.x, d.t = x.data_ptr(), t.data_ptr()
    if cond is not None:
        d.cond_dim = cond.shape[1]
        d.cond = cond.data_ptr()
    return d


class _MlpFn(torch.autograd.Function):
    """a(x, cond, t) with a backward on the library's own kernels (`dmip_mlp_forward_stash` / `dmip_mlp_backward`,
    include/dmip.h): what `loss.backward()` of a user-written loss runs instead of torch's addmm / tanh chain.
    First-order only — the reference's double-backward losses (losses.py:14-26) are the fused PINNLoss / DSM_PDELoss."""

    @staticmethod
    def forward(ctx, net, x, cond, t, *params):
        L = _bind_grad()
        keep = [x, cond, t]
        d = _grad_desc(net, x, cond, t, keep)
        out = torch.empty(x.shape[0], net.output_dim, device=x.device, dtype=torch.float32)
        d.out = out.data_ptr()
        nbytes = L.dmip_mlp_grad_workspace_bytes(C.byref(d))
        ws = torch.empty(nbytes + 1024, dtype=torch.uint8, device=x.device)
        d.workspace = (ws.data_ptr() + 1023) & ~1023
        d.workspace_bytes = nbytes
        with torch.cuda.device(x.device):
            _lib.check(L.dmip_mlp_forward_stash(C.byref(d), _lib.stream_ptr()))
        ctx.d, ctx.keep, ctx.ws = d, keep, ws
        ctx.shapes = [p.shape for p in params]
        ctx.dims = (x.shape[1], 0 if cond is None else cond.shape[1])
        net.last_launch_count = L.dmip_last_launch_count()
        return out

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, gout):
        L = _bind_grad()
        d = ctx.d
        gout = gout.contiguous().float()
        dev = gout.device
        flat = torch.empty(L.dmip_loss_grad_floats(C.byref(d.net)), device=dev, dtype=torch.float32)
        need_in = any(ctx.needs_input_grad[1:4])
        gin = torch.empty(gout.shape[0], d.net.in_dim, device=dev, dtype=torch.float32) if need_in else None
        d.grad_out, d.grad_params = gout.data_ptr(), flat.data_ptr()
        d.grad_in = gin.data_ptr() if need_in else None
        with torch.cuda.device(dev):
            _lib.check(L.dmip_mlp_backward(C.byref(d), _lib.stream_ptr()))
        grads, off = [], 0
        for shp in ctx.shapes:
            grads.append(flat[off:off + shp.numel()].view(shp))
            off += shp.numel()
        xd, cd = ctx.dims
        gx = gin[:, :xd] if need_in and ctx.needs_input_grad[1] else None
        gc = gin[:, xd:xd + cd] if need_in and cd and ctx.needs_input_grad[2] else None
        gt = gin[:, xd + cd:].reshape(ctx.keep[2].shape) if need_in and ctx.needs_input_grad[3] else None
        return (None, gx, gc, gt, *grads)


class _ScoreMLP(nn.Sequential):
    """[in] -> hidden... -> [out] tanh MLP with the reference's module naming."""

    def __init__(self, input_dim, output_dim, hidden_layers, activation):
        super().__init__()
        self.input_dim = input_dim
        self.output_dim = output_dim
        self.hidden_layers = hidden_layers
        widths = [input_dim] + list(hidden_layers)
        self.add_module('0', nn.Linear(widths[0], widths[1]))
        self.add_module('1', activation)
        self.add_module('act', activation)          # second registration of the same module: double tanh (Q1)
        for fan_in, fan_out in zip(widths[1:-1], widths[2:]):
            self.add_module(str(len(self)), nn.Linear(fan_in, fan_out))
            self.add_module(str(len(self)), activation)
        self.add_module(str(len(self)), nn.Linear(widths[-1], output_dim))
        self._packed = _lib.PackedNet()
        self.precision = 'bf16'
        self.l0_split = 4                    # layer-0 operand: one f16 part (1: bf16, 2: bf16 hi + lo, 3: + W0 split)

    def _fused(self, x, cond, t):
        L = _lib.require_gpu()
        n = x.shape[0]
        keep = []
        d = _lib.DmipForward()
        d.net = _lib.mlp_desc(self, keep)
        prec = _lib.precision_code(self.precision)
        if prec == _lib.PREC_BF16 and not _lib.tc_supported(self, None, self.l0_split):
            prec = _lib.PREC_F32                     # other widths / too deep a layer 0: fp32 FFMA kernels (still CUDA, never CPU)
        self.last_precision = 'bf16' if prec == _lib.PREC_BF16 else 'fp32'   # the path that actually ran
        d.precision = prec
        d.l0_split = self.l0_split
        d.n = n
        x = x.detach().float().contiguous()
        t = t.detach().float().reshape(-1).contiguous()
        assert t.numel() == n, 'Input Tensor is expected to be 2D with shape (batch_size, xdim+ydim+1)'
        d.x_dim = x.shape[1]
        d.x = x.data_ptr()
        d.t = t.data_ptr()
        if cond is not None and cond.numel() > 0:
            cond = cond.detach().float().contiguous()
            d.cond_dim = cond.shape[1]
            d.cond = cond.data_ptr()
        out = torch.empty(n, self.output_dim, device=x.device, dtype=torch.float32)
        d.out = out.data_ptr()
        with torch.cuda.device(x.device):            # launch on the tensors' device and its current stream
            if prec == _lib.PREC_BF16:
                d.packed = self._packed.get(self, d.net.in_dim, self.output_dim, self.l0_split).data_ptr()
            else:
                ws = torch.empty(max(L.dmip_forward_workspace_bytes(C.byref(d)), 16), dtype=torch.uint8, device=x.device)
                keep.append(ws)
                d.workspace = ws.data_ptr()
                d.workspace_bytes = ws.numel()
            _lib.check(L.dmip_mlp_forward(C.byref(d), _lib.stream_ptr()))
        return out

    def _autograd_fused(self, x, cond, t):
        """Autograd through the library's own kernels (forward with stash + backward), or None when the net's shape is
        outside what they serve."""
        L = _bind_grad()
        if x.shape[0] == 0:
            return None
        xx = x.float().contiguous()
        cc = None if cond is None or cond.numel() == 0 else cond.float().contiguous()
        tt = t.float().reshape(-1).contiguous()
        assert tt.numel() == xx.shape[0], 'Input Tensor is expected to be 2D with shape (batch_size, xdim+ydim+1)'
        probe = _grad_desc(self, xx.detach(), None if cc is None else cc.detach(), tt.detach(), [])
        if L.dmip_mlp_grad_workspace_bytes(C.byref(probe)) == 0:
            return None
        params = [p for lin in _lib.linear_layers(self) for p in (lin.weight, lin.bias)]
        return _MlpFn.apply(self, xx, cc, tt, *params)

    def invalidate_packed(self):
        """Drop the cached tcgen05 weight image (needed only after writes through `p.data`, see _lib.PackedNet)."""
        self._packed.invalidate()

    def _dispatch(self, x, cond, t):
        params_need_grad = torch.is_grad_enabled() and (
            x.requires_grad or t.requires_grad or (cond is not None and cond.requires_grad)
            or any(p.requires_grad for p in self.parameters()))
        if x.is_cuda and not params_need_grad:
            return self._fused(x, cond, t)
        if not x.is_cuda:
            raise RuntimeError("dmip score nets run on CUDA (sm_100a) only: there is no CPU fallback")
        fused = self._autograd_fused(x, cond, t)
        if fused is not None:
            return fused
        if not _WARNED_EAGER[0]:
            _WARNED_EAGER[0] = True
            warnings.warn("dmip: autograd is recording through MLP.forward of a net the tcgen05 kernels do not serve (they take "
                          "[in <= 64] -> 512 -> 512 -> 512 -> [out <= 64]) — this call runs the plain torch module chain "
                          "(cuBLAS addmm), not the fused sm_100a kernels.", RuntimeWarning, stacklevel=3)
        parts = [x] + ([cond] if cond is not None and cond.numel() > 0 else []) + [t.view(len(x), 1)]
        inp = torch.cat(parts, dim=1)
        assert inp.ndim == 2, 'Input Tensor is expected to be 2D with shape (batch_size, xdim+ydim+1)'
        return nn.Sequential.forward(self, inp)


class MLP(_ScoreMLP):
    """a(x, y, t) on cat[x, y, t]; `y` may be the empty tensor (CDiffE losses)."""

    def forward(self, x, y, t):
        return self._dispatch(x, y, t)


class MLP2(_ScoreMLP):
    """s(x, t) on cat[x, t] — the DPS prior net."""

    def forward(self, x, t):
        return self._dispatch(x, None, t)


class PosteriorScore(nn.Module):
    """g(t, x) * (prior_net(x, t) + likelihood_net(x, y, t)) — the DPS drift."""

    def __init__(self, prior_net, likelihood_net, forward_process):
        super().__init__()
        self.prior_net = prior_net
        self.likelihood_net = likelihood_net
        self.forward_sde = forward_process

    def forward(self, x, y, t):
        return self.forward_sde.g(t, x) * (self.prior_net(x, t) + self.likelihood_net(x, y, t))
