"""Helpers for the -m gpu parity tests: call the C ABI on torch CUDA tensors."""
import ctypes as C

import numpy as np
import torch

from basi_b200 import _lib
from basi_b200.engine import Act

DEV = "cuda:0"


def stream():
    return torch.cuda.current_stream().cuda_stream


def act(x_np, dtype=torch.float32):
    """numpy NHWC -> device Act"""
    return Act(torch.from_numpy(np.ascontiguousarray(x_np)).to(DEV).to(dtype).contiguous())


def empty_act(shape, dtype=torch.float32, fill=None):
    t = torch.zeros(*shape, dtype=dtype, device=DEV)
    if fill is not None:
        t.fill_(fill)
    return Act(t)


def dev(x_np, dtype=None):
    t = torch.from_numpy(np.ascontiguousarray(x_np)).to(DEV)
    return t if dtype is None else t.to(dtype)


def call(name, *args):
    return _lib.call(name, *args, stream())


def host(a):
    t = a.t if isinstance(a, Act) else a
    torch.cuda.synchronize()
    return t.float().cpu().numpy() if t.dtype == torch.bfloat16 else t.cpu().numpy()


def bf16_round(x_np):
    return torch.from_numpy(np.ascontiguousarray(x_np, dtype=np.float32)).to(torch.bfloat16).float().numpy()


def rel_err(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-30))


def rel_l2(a, b):
    """norm-wise relative error: one ReLU-mask tie (|pre-activation| below float32 resolution) may flip a single
    element of a gradient by its full value; that must not fail a 2-million-element comparison"""
    a, b = np.asarray(a, dtype=np.float64).reshape(-1), np.asarray(b, dtype=np.float64).reshape(-1)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))
