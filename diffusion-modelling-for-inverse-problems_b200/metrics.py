"""Evaluation metrics computed next to the sampler: the reference's 75-bin sample histograms and the
histogram KL it reports (`evaluate` in main_diffusion_linear.py:84-117 and main_diffusion_scatterometry.py:72-101),
on the GPU where the samples already are — `model(y, ..., return_tensor=True)` feeds them without a device->host copy.

    acc = HistogramKL(bins=(75, 75), ranges=((-3.5, 3.5), (-3.5, 3.5)))
    for repeat in range(10):
        acc.add(x_true, x_pred)          # CUDA tensors (n, d) — the reference's  hist_sum += np.histogramdd(...)
    kl2 = acc.kl()                       # sum(rel_entr(hist_true, hist_diffusion)) with the reference's normalisation
"""
import ctypes as C

import numpy as np
import torch

from . import _lib


class DmipHistogram(C.Structure):
    _fields_ = [("dim", C.c_int32), ("bins", C.c_int32 * 4), ("edges", C.c_void_p * 4), ("n", C.c_int64),
                ("x", C.c_void_p), ("counts", C.c_void_p)]


def _bind():
    L = _lib.require_gpu()
    if not getattr(L, "_metrics_bound", False):
        L.dmip_histogramdd.restype = C.c_int
        L.dmip_histogramdd.argtypes = [C.POINTER(DmipHistogram), C.c_void_p]
        L.dmip_hist_kl.restype = C.c_int
        L.dmip_hist_kl.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_double, C.c_void_p, C.c_void_p]
        L._metrics_bound = True
    return L


def histogramdd(x, bins, ranges, out=None):
    """np.histogramdd(x, bins=bins, range=ranges)[0] for a CUDA tensor x (n, d), bit-exact, as an int64 CUDA tensor of
    shape `bins`; with `out` the counts are accumulated into it."""
    L = _bind()
    if not x.is_cuda:
        raise RuntimeError("dmip metrics run on CUDA (sm_100a) only: there is no CPU fallback")
    x = x.detach().float().contiguous()
    n, dim = x.shape
    if len(bins) != dim or len(ranges) != dim:
        raise ValueError("The dimension of bins must be equal to the dimension of the sample x.")
    if out is None:
        out = torch.zeros(tuple(int(b) for b in bins), dtype=torch.int64, device=x.device)
    d = DmipHistogram()
    d.dim, d.n = dim, n
    keep = []
    for k in range(dim):
        e = torch.from_numpy(np.linspace(ranges[k][0], ranges[k][1], int(bins[k]) + 1)).to(x.device)   # float64, as numpy
        keep.append(e)
        d.bins[k] = int(bins[k])
        d.edges[k] = e.data_ptr()
    d.x, d.counts = x.data_ptr(), out.data_ptr()
    with torch.cuda.device(x.device):
        _lib.check(L.dmip_histogramdd(C.byref(d), _lib.stream_ptr()))
    return out


def hist_kl(hist_p, hist_q, epsilon=1e-10):
    """sum(scipy.special.rel_entr(p, q)) after the reference's normalisation (h / sum, + epsilon, / sum)."""
    L = _bind()
    assert hist_p.shape == hist_q.shape and hist_p.dtype == torch.int64 and hist_q.dtype == torch.int64
    out = torch.empty(3, dtype=torch.float64, device=hist_p.device)
    with torch.cuda.device(hist_p.device):
        _lib.check(L.dmip_hist_kl(hist_p.contiguous().data_ptr(), hist_q.contiguous().data_ptr(), hist_p.numel(),
                                  float(epsilon), out.data_ptr(), _lib.stream_ptr()))
    return out[0]


class HistogramKL:
    """Running pair of histograms over repeats and their KL — the inner loop of the reference's `evaluate`."""

    def __init__(self, bins, ranges, epsilon=1e-10):
        self.bins, self.ranges, self.epsilon = tuple(bins), tuple(ranges), epsilon
        self.hist_true = None
        self.hist_model = None

    def add(self, x_true, x_model):
        self.hist_true = histogramdd(x_true, self.bins, self.ranges, self.hist_true)
        self.hist_model = histogramdd(x_model, self.bins, self.ranges, self.hist_model)

    def kl(self, reverse=False):
        p, q = (self.hist_model, self.hist_true) if reverse else (self.hist_true, self.hist_model)
        return float(hist_kl(p, q, self.epsilon).item())
