"""Drop-in for the reference's `nets.py` score networks (MLP :17-35, MLP2 :37-57, PosteriorScore :143-157).

The modules keep the reference's parameter layout — `state_dict()` keys `0,3,5,7.{weight,bias}` —
and its numerics: tanh is applied twice after the first Linear, because the reference registers its activation a
second time under the name 'act' and `nn.Sequential.forward` walks every registered module.  That is
reproduced here on purpose so reference checkpoints (`current_model.pt`, `diffusion.pt`) load and behave identically.

`forward` on CUDA tensors without autograd runs the fused kernels of libdmip_sm100.so (`dmip_mlp_forward`);
when autograd is recording it runs the plain module chain so `loss.backward()` of user-written losses keeps working.
`GaussianFourierProjection` / `TemporalMLP*` of the reference are unused by every model (nets.py:62-63) and omitted.
"""
import ctypes as C
import warnings

import torch
from torch import nn

from . import _lib

_WARNED_EAGER = [False]


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
        self.l0_split = 2

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
        if not _WARNED_EAGER[0]:
            _WARNED_EAGER[0] = True
            warnings.warn("dmip: autograd is recording through MLP.forward — this call runs the plain torch module chain "
                          "(cuBLAS addmm), not the fused sm_100a kernels.  The fused kernels serve the no-grad calls "
                          "(sampler, evaluation) and the fused losses (DSMLoss via train_epoch, PINNLoss, DSM_PDELoss, "
                          "PosteriorLoss), which carry their own backward; wrap inference in torch.no_grad().",
                          RuntimeWarning, stacklevel=3)
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
