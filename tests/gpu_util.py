"""Helpers shared by the GPU parity tests (they call the product ONLY through the C-ABI)."""
import os

import numpy as np
import torch

from wct_tf_b200 import _capi

OUT_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")


def lib():
    return _capi.load()


def stream():
    return torch.cuda.current_stream().cuda_stream


def dev(a, dtype=None):
    t = torch.from_numpy(np.ascontiguousarray(a))
    if dtype is not None:
        t = t.to(dtype)
    return t.cuda()


def act_from_numpy(x_nhwc):
    """fp32 NHWC numpy -> (device SPF16 buffer, shape)."""
    x = np.ascontiguousarray(x_nhwc, dtype=np.float32)
    n, h, w, c = x.shape
    src = dev(x)
    buf = torch.empty(lib().wctb200_act_bytes(n, h, w, c), dtype=torch.uint8, device="cuda")
    _capi.check(lib().wctb200_act_from_f32(src.data_ptr(), n, h, w, c, buf.data_ptr(), stream()))
    return buf


def act_alloc(n, h, w, c, poison=True):
    buf = torch.empty(lib().wctb200_act_bytes(n, h, w, c), dtype=torch.uint8, device="cuda")
    if poison:
        buf.view(torch.float16).fill_(float("nan"))   # anything left unwritten shows up as NaN
    return buf


def act_to_numpy(buf, n, h, w, c):
    out = torch.empty((n, h, w, c), dtype=torch.float32, device="cuda")
    _capi.check(lib().wctb200_act_to_f32(buf.data_ptr(), n, h, w, c, out.data_ptr(), stream()))
    torch.cuda.synchronize()
    return out.cpu().numpy()


def act_raw_padded(buf, n, h, w, c):
    """hi+lo in float64 over the whole padded plane [n, h+2, w+2, c]."""
    planes = buf.view(torch.float16).view(2, n, h + 2, w + 2, c).cpu().numpy().astype(np.float64)
    return planes[0] + planes[1]


def check_device():
    _capi.check(lib().wctb200_check_device(stream()))


def split_repr(x):
    """The value SPF16 actually stores for fp32 x (hi + lo as fp16 pair)."""
    x = np.asarray(x, dtype=np.float32)
    hi = x.astype(np.float16)
    lo = (x - hi.astype(np.float32)).astype(np.float16)
    return hi.astype(np.float64) + lo.astype(np.float64)


def dump(name, **arrays):
    os.makedirs(OUT_DIR, exist_ok=True)
    np.savez_compressed(os.path.join(OUT_DIR, name + ".npz"), **arrays)
