"""shared helpers for the parity tests"""
import numpy as np
import torch


def rel_l2(a, b):
    a = np.asarray(a, np.float64).ravel()
    b = np.asarray(b, np.float64).ravel()
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


def gpu(x, dtype=torch.float32):
    return torch.as_tensor(np.ascontiguousarray(x)).to(device="cuda", dtype=dtype).contiguous()


def cpu(t):
    torch.cuda.synchronize()
    return t.detach().float().cpu().numpy()


def t(x, dtype=torch.float32, grad=False):
    r = torch.as_tensor(np.ascontiguousarray(x)).to(dtype)
    if grad:
        r.requires_grad_(True)
    return r
