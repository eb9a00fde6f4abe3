"""Drop-in for the reference's `losses.py` — filled in by the fused loss kernels (see dmip_loss.cu)."""
import torch
from torch import nn


class DSMLoss(nn.Module):
    def __init__(self):
        super().__init__()
        self.name = 'DSMLoss'

    def forward(self, s, std, target):
        return ((s * std + target) ** 2).view(s.shape[0], -1).sum(1, keepdim=False) / 2


class PosteriorLoss(nn.Module):
    def __init__(self, forward_model, a, b, lam):
        super().__init__()
        self.name = 'PosteriorLoss'
        self.forward_model, self.a, self.b, self.lam = forward_model, a, b, lam


def fused_train_step(model, loss_fn, x, y, t):
    raise NotImplementedError("fused loss kernels not built yet")
