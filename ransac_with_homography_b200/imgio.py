"""Device-side image I/O around the hot path: the GPU counterparts of the reference's
``cv2.imread`` (app.py:122,154,158; ransac.py:291-294) and ``cv2.imwrite`` (app.py:193-195; ransac.py:301).

``imread`` / ``imwrite`` keep cv2's call shapes (BGR numpy arrays, quality 95) for drop-in use; ``decode`` / ``encode``
work on device tensors so that a frame goes JPEG bytes -> HBM -> warp / stitch -> HBM -> JPEG bytes and only the
compressed bytes cross the host link.  The codec is nvJPEG behind ``b200geo_jpeg_*`` (include/b200geo.h (7)); there is
no CPU fallback.  Only JPEG is handled (the reference's fixtures and its ``imwrite(filename + '.jpg')``); other
extensions raise ValueError.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import _cabi

BGR, RGB = 0, 1
_ORDER = {"bgr": BGR, "rgb": RGB}


def _bytes_array(data) -> np.ndarray:
    if isinstance(data, (bytes, bytearray, memoryview)):
        return np.frombuffer(data, dtype=np.uint8)
    a = np.ascontiguousarray(data, dtype=np.uint8).reshape(-1)
    return a


def info(data, device: int | None = None):
    """(height, width, components) of a JPEG byte string."""
    a = _bytes_array(data)
    ctx = _cabi.context(device)
    h, w, c = C.c_int(), C.c_int(), C.c_int()
    _cabi.check(ctx.lib.b200geo_jpeg_info(ctx.handle, _cabi.ptr(a), a.size, C.byref(h), C.byref(w), C.byref(c)))
    return h.value, w.value, c.value


def imdecode(data, order: str = "bgr", device: int | None = None) -> np.ndarray:
    """cv2.imdecode(data, cv2.IMREAD_COLOR) on the GPU: (h, w, 3) uint8 on the host, BGR unless order='rgb'."""
    a = _bytes_array(data)
    h, w, _ = info(a, device)
    out = np.empty((h, w, 3), np.uint8)
    ctx = _cabi.context(device)
    _cabi.check(ctx.lib.b200geo_jpeg_decode(ctx.handle, _cabi.ptr(a), a.size, _ORDER[order], _cabi.ptr(out), 3 * w, 0, None))
    return out


def imread(filename, order: str = "bgr", device: int | None = None):
    """cv2.imread(filename) for JPEG files; None if the file cannot be read (as cv2.imread)."""
    try:
        with open(filename, "rb") as f:
            data = f.read()
    except OSError:
        return None
    if not data:
        return None
    return imdecode(data, order, device)


def imencode(img: np.ndarray, quality: int = 95, order: str = "bgr", device: int | None = None) -> bytes:
    """cv2.imencode('.jpg', img)[1].tobytes() on the GPU (quality 95, 4:2:0, as cv2's defaults)."""
    assert img.ndim == 3 and img.shape[2] == 3 and img.dtype == np.uint8
    src = np.ascontiguousarray(img)
    h, w = src.shape[:2]
    cap = h * w * 3 + 4096
    out = np.empty(cap, np.uint8)
    n = C.c_size_t()
    ctx = _cabi.context(device)
    _cabi.check(ctx.lib.b200geo_jpeg_encode(ctx.handle, _cabi.ptr(src), h, w, 3 * w, _ORDER[order], 0, int(quality), _cabi.ptr(out), cap,
                                            C.byref(n), None))
    return out[:n.value].tobytes()


def imwrite(filename, img: np.ndarray, quality: int = 95, order: str = "bgr", device: int | None = None) -> bool:
    """cv2.imwrite(filename, img) for .jpg / .jpeg names."""
    if os.path.splitext(str(filename))[1].lower() not in (".jpg", ".jpeg"):
        raise ValueError("only JPEG is encoded on the device")
    data = imencode(img, quality, order, device)
    with open(filename, "wb") as f:
        f.write(data)
    return True


def decode(data, order: str = "rgb", out=None):
    """JPEG bytes -> (h, w, 3) uint8 CUDA tensor, asynchronous on the current torch stream after the entropy decode."""
    import torch
    a = _bytes_array(data)
    dev = out.device.index if out is not None else torch.cuda.current_device()
    h, w, _ = info(a, dev)
    if out is None:
        out = torch.empty((h, w, 3), dtype=torch.uint8, device=f"cuda:{dev}")
    assert out.is_cuda and out.dtype == torch.uint8 and tuple(out.shape) == (h, w, 3) and out.stride(2) == 1 and out.stride(1) == 3
    ctx = _cabi.context(dev)
    st = _cabi._vp(torch.cuda.current_stream(out.device).cuda_stream)
    _cabi.check(ctx.lib.b200geo_jpeg_decode(ctx.handle, _cabi.ptr(a), a.size, _ORDER[order], _cabi._vp(out.data_ptr()), out.stride(0), 1, st))
    return out


def encode(img, quality: int = 95, order: str = "rgb") -> bytes:
    """(h, w, 3) uint8 CUDA tensor -> JPEG bytes (synchronises the current stream)."""
    import torch
    assert img.is_cuda and img.dtype == torch.uint8 and img.dim() == 3 and img.shape[2] == 3 and img.stride(2) == 1 and img.stride(1) == 3
    h, w = int(img.shape[0]), int(img.shape[1])
    cap = h * w * 3 + 4096
    out = np.empty(cap, np.uint8)
    n = C.c_size_t()
    ctx = _cabi.context(img.device.index)
    st = _cabi._vp(torch.cuda.current_stream(img.device).cuda_stream)
    _cabi.check(ctx.lib.b200geo_jpeg_encode(ctx.handle, _cabi._vp(img.data_ptr()), h, w, img.stride(0), _ORDER[order], 1, int(quality),
                                            _cabi.ptr(out), cap, C.byref(n), st))
    return out[:n.value].tobytes()
