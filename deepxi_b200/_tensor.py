"""Tensor plumbing shared by the API mirror modules: numpy / torch in, same kind out, CUDA in between."""
import numpy as np
import torch


def device():
    if not torch.cuda.is_available():
        raise RuntimeError('deepxi_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback')
    return torch.device('cuda', torch.cuda.current_device())


def to_dev(x, dtype=torch.float32):
    """Returns (contiguous CUDA tensor of `dtype`, was_numpy)."""
    was_numpy = not isinstance(x, torch.Tensor)
    if was_numpy:
        x = torch.from_numpy(np.ascontiguousarray(np.asarray(x)))
    if x.dtype != dtype:
        x = x.to(dtype)
    if not x.is_cuda:
        x = x.to(device(), non_blocking=True)
    return x.contiguous(), was_numpy


def ret(t, was_numpy):
    return t.cpu().numpy() if was_numpy else t
