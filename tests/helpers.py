"""Shared helpers for the parity tests."""
import numpy as np
import torch


def rel_l2(a, b):
    a = torch.as_tensor(a).detach().cpu()
    b = torch.as_tensor(b).detach().cpu()
    if a.is_complex():
        a, b = torch.view_as_real(a), torch.view_as_real(b)
    a, b = a.double(), b.double()
    return float((a - b).pow(2).sum().sqrt() / b.pow(2).sum().sqrt().clamp_min(1e-30))


def nchw_to_ntfc(x):
    """reference layout [B,C,F,T] -> kernel layout [B,T,F,C] (contiguous)."""
    return x.permute(0, 3, 2, 1).contiguous()


def ntfc_to_nchw(x):
    return x.permute(0, 3, 2, 1).contiguous()


def load_npz(path):
    with np.load(path) as z:
        return {k: z[k] for k in z.files}
