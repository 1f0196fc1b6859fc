"""ctypes binding of libb200geo.so (C ABI declared in include/b200geo.h).

This is the only place the package touches native code.  There is NO CPU fallback: if the shared
library is missing, or no sm_100 GPU is visible, the first compute call raises.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import threading

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("B200GEO_LIB") or os.path.join(_HERE, "libb200geo.so")
CSRC = os.path.join(_HERE, "csrc")

OK, ERR_CUDA, ERR_ARG, ERR_NODEVICE, ERR_SINGULAR, ERR_NOMODEL = 0, -1, -2, -3, -4, -5
FWD, BACKWARD, REPROJ = 0, 1, 2
NN, BILINEAR = 0, 1
U8, F32, F64 = 0, 1, 2
METHODS = {"fwd": FWD, "backward": BACKWARD, "reproj": REPROJ}
SAMPLERS = {"nn": NN, "bilinear": BILINEAR}


class B200GeoError(RuntimeError):
    def __init__(self, code, text):
        super().__init__(f"b200geo error {code}: {text}")
        self.code = code


class WarpJob(C.Structure):
    """struct b200geo_warp_job (include/b200geo.h)"""
    _fields_ = [("invH", C.c_double * 9),
                ("x0", C.c_double), ("dx", C.c_double), ("x_last", C.c_double),
                ("y0", C.c_double), ("dy", C.c_double), ("y_last", C.c_double),
                ("ow", C.c_int32), ("oh", C.c_int32), ("bw", C.c_int32), ("bh", C.c_int32),
                ("iw", C.c_int32), ("ih", C.c_int32), ("reserved0", C.c_int32), ("reserved1", C.c_int32),
                ("src_offset", C.c_int64), ("src_pitch", C.c_int64),
                ("dst_offset", C.c_int64), ("dst_pitch", C.c_int64)]


WARP_JOB_DTYPE = np.dtype([("invH", np.float64, 9), ("x0", np.float64), ("dx", np.float64), ("x_last", np.float64),
                           ("y0", np.float64), ("dy", np.float64), ("y_last", np.float64),
                           ("ow", np.int32), ("oh", np.int32), ("bw", np.int32), ("bh", np.int32),
                           ("iw", np.int32), ("ih", np.int32), ("reserved0", np.int32), ("reserved1", np.int32),
                           ("src_offset", np.int64), ("src_pitch", np.int64),
                           ("dst_offset", np.int64), ("dst_pitch", np.int64)])
assert WARP_JOB_DTYPE.itemsize == C.sizeof(WarpJob) == 184

class Compose(C.Structure):
    """struct b200geo_compose (include/b200geo.h)"""
    _fields_ = [("fw", C.c_int32), ("fh", C.c_int32), ("tsx", C.c_int32), ("tsy", C.c_int32),
                ("qsx", C.c_int32), ("qsy", C.c_int32), ("qw", C.c_int32), ("qh", C.c_int32),
                ("q_pitch", C.c_int64), ("canvas_pitch", C.c_int64),
                ("alpha_t", C.c_double), ("alpha_q", C.c_double), ("alpha_bg", C.c_double),
                ("blend", C.c_int32), ("reserved", C.c_int32)]


class SrcLayout(C.Structure):
    """struct b200geo_src_layout (include/b200geo.h)"""
    _fields_ = [("pitch", C.c_int64), ("rows", C.c_int64), ("image_rows", C.c_int64), ("row_bytes", C.c_int64)]


_vp, _i, _f, _d, _sz = C.c_void_p, C.c_int, C.c_float, C.c_double, C.c_size_t

# name -> (restype, argtypes); every symbol include/b200geo.h declares
SIGNATURES = {
    "b200geo_abi_version": (_i, []),
    "b200geo_last_error": (C.c_char_p, []),
    "b200geo_create": (_i, [_i, C.POINTER(_vp)]),
    "b200geo_destroy": (_i, [_vp]),
    "b200geo_device_info": (_i, [_vp, C.POINTER(_i), C.POINTER(_i), C.POINTER(_i), C.POINTER(_sz)]),
    "b200geo_launch_count": (C.c_int64, [_vp]),
    "b200geo_sync": (_i, [_vp, _vp]),
    "b200geo_stats": (_i, [_vp, C.POINTER(C.c_ulonglong), _i]),
    "b200geo_set_option": (_i, [_vp, _i, _i]),
    "b200geo_warp_stats": (_i, [_vp, C.POINTER(C.c_ulonglong), _i]),
    "b200geo_dlt4_dev": (_i, [_vp, _vp, _i, _vp, _i, _vp, _vp, _vp, _vp]),
    "b200geo_dlt4_host": (_i, [_vp, _vp, _i, _vp, _i, _vp, _vp, _vp]),
    "b200geo_homography4_host": (_i, [_vp, _vp, _vp, _vp]),
    "b200geo_score_dev": (_i, [_vp, _vp, _i, _vp, _vp, _vp, _i, _f, _i, _vp, _vp]),
    "b200geo_score_host": (_i, [_vp, _vp, _i, _vp, _vp, _i, _f, _i, _vp]),
    "b200geo_loss_host": (_i, [_vp, _vp, _i, _vp, _vp, _i, _vp]),
    "b200geo_project_host": (_i, [_vp, _vp, _vp, _i, _vp]),
    "b200geo_dist_host": (_i, [_vp, _vp, _vp, _i, _i, _vp]),
    "b200geo_ransac_search_dev": (_i, [_vp, _vp, _i, _vp, _i, _i, _f, _d, _i, _vp, _i, _vp, _vp]),
    "b200geo_ransac_finish_dev": (_i, [_vp, _vp, _i, _vp, _vp, _f, _i, _vp, _vp, _vp]),
    "b200geo_ransac_search_shard_dev": (_i, [_vp, _vp, _i, _vp, _i, _i, _f, _d, _i, _vp, _vp, _vp]),
    "b200geo_ransac_finish_recs_dev": (_i, [_vp, _vp, _i, _vp, _i, _f, _i, _vp, _vp, _vp]),
    "b200geo_ransac_host": (_i, [_vp, _vp, _i, _vp, _i, _f, _d, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "b200geo_ransac_batched_dev": (_i, [_vp, _vp, _i, _vp, _i, _i, _f, _d, _i, _vp, _vp, _vp]),
    "b200geo_refit_host": (_i, [_vp, _vp, _i, _vp, _vp]),
    "b200geo_homography_svd_host": (_i, [_vp, _vp, _i, _vp]),
    "b200geo_warp_dev": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _vp]),
    "b200geo_warp_jobs_dev": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _vp]),
    "b200geo_warp_batch_dev": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _vp, _vp]),
    "b200geo_warp_batch_host": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i]),
    "b200geo_warp_host": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i]),
    "b200geo_sample_host": (_i, [_vp, _vp, _i, _i, _i, _i, _vp, _vp, C.c_longlong, _i, _i, _i, _i, _vp]),
    "b200geo_match_hamming_host": (_i, [_vp, _vp, _i, _vp, _i, _i, _i, _vp, _vp]),
    "b200geo_remap_linear_host": (_i, [_vp, _vp, _i, _i, _i, _vp, _vp, _i, _i, _vp]),
    "b200geo_orb_host": (_i, [_vp, _vp, _i, _i, C.c_longlong, _i, _i, _vp, _vp, _i, _vp]),
    "b200geo_peer_export": (_i, [_vp, _i, _vp]),
    "b200geo_peer_open": (_i, [_vp, _i, _i, _vp]),
    "b200geo_peer_close": (_i, [_vp]),
    "b200geo_ransac_exchange_peer_dev": (_i, [_vp, _vp, _vp, _vp]),
    "b200geo_jpeg_info": (_i, [_vp, _vp, C.c_size_t, _vp, _vp, _vp]),
    "b200geo_jpeg_decode": (_i, [_vp, _vp, C.c_size_t, _i, _vp, C.c_longlong, _i, _vp]),
    "b200geo_jpeg_encode": (_i, [_vp, _vp, _i, _i, C.c_longlong, _i, _i, _i, _vp, C.c_size_t, _vp, _vp]),
    "b200geo_stitch_jobs_dev": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _vp, _vp, _vp]),
    "b200geo_stitch_compose_dev": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "b200geo_stitch_compose_host": (_i, [_vp, _vp, _vp, _vp, _vp, _vp]),
}

_lib = None
_lock = threading.Lock()
_contexts = {}


def build(force: bool = False) -> str:
    """Compile libb200geo.so in-tree with nvcc for sm_100a (needs no GPU)."""
    args = ["make", "-C", CSRC, "-s"]
    if force:
        args.append("-B")
    subprocess.run(args, check=True)
    return LIB_PATH


def load() -> C.CDLL:
    """dlopen libb200geo.so and bind every declared symbol.  Raises if the library is absent."""
    global _lib
    if _lib is None:
        with _lock:
            if _lib is None:
                if not os.path.exists(LIB_PATH):
                    raise ImportError(
                        f"{LIB_PATH} not found: build it with `make -C {CSRC}` (or __graft_entry__.build()). "
                        "ransac_with_homography_b200 has no CPU fallback.")
                lib = C.CDLL(LIB_PATH)
                for name, (res, args) in SIGNATURES.items():
                    fn = getattr(lib, name)     # AttributeError if the symbol is missing
                    fn.restype, fn.argtypes = res, args
                if lib.b200geo_abi_version() != 1:
                    raise ImportError("libb200geo.so ABI version mismatch; rebuild it")
                _lib = lib
    return _lib


def last_error() -> str:
    return load().b200geo_last_error().decode("utf-8", "replace")


def check(rc: int) -> None:
    if rc != OK:
        text = last_error()
        if rc == ERR_SINGULAR:
            raise np.linalg.LinAlgError(text or "Singular matrix")
        raise B200GeoError(rc, text)


class Context:
    """One b200geo_ctx (one GPU)."""

    def __init__(self, device: int = 0):
        lib = load()
        h = _vp()
        check(lib.b200geo_create(int(device), C.byref(h)))
        self.lib, self.handle, self.device = lib, h, int(device)
        sm, cc, khz, mem = _i(), _i(), _i(), _sz()
        check(lib.b200geo_device_info(h, C.byref(sm), C.byref(cc), C.byref(khz), C.byref(mem)))
        self.sm_count, self.cc, self.clock_khz, self.mem_bytes = sm.value, cc.value, khz.value, mem.value

    def launch_count(self) -> int:
        return int(self.lib.b200geo_launch_count(self.handle))

    def exact_path_pairs(self, reset: bool = False) -> int:
        """Pairs scored by the reference-order chain because they fell inside the threshold shell."""
        v = C.c_ulonglong(0)
        check(self.lib.b200geo_stats(self.handle, C.byref(v), int(reset)))
        return int(v.value)

    def force_exact(self, on: bool) -> None:
        check(self.lib.b200geo_set_option(self.handle, 1, int(on)))

    def warp_exact(self, on: bool) -> None:
        """Warp uint8 RGB images with the float64 kernel only (bit-exact bilinear) instead of the fast path."""
        check(self.lib.b200geo_set_option(self.handle, 2, int(on)))

    def warp_variant(self, v: int) -> None:
        """0 = persistent TMA-pipelined fast warp kernel (default), 1 = one CTA per tile."""
        check(self.lib.b200geo_set_option(self.handle, 3, int(v)))

    def warp_exact_pixels(self, reset: bool = False) -> int:
        """Output pixels the fast warp kernel evaluated in float64 (band / slow tiles)."""
        v = C.c_ulonglong(0)
        check(self.lib.b200geo_warp_stats(self.handle, C.byref(v), int(reset)))
        return int(v.value)

    def sync(self, stream=None) -> None:
        check(self.lib.b200geo_sync(self.handle, _vp(stream or 0)))

    def close(self) -> None:
        if self.handle:
            self.lib.b200geo_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def context(device: int | None = None) -> Context:
    """Process-wide context for `device` (default: $B200GEO_DEVICE, else $LOCAL_RANK, else 0)."""
    if device is None:
        device = int(os.environ.get("B200GEO_DEVICE", os.environ.get("LOCAL_RANK", "0")))
    with _lock:
        ctx = _contexts.get(device)
    if ctx is None:
        ctx = Context(device)
        with _lock:
            _contexts.setdefault(device, ctx)
            ctx = _contexts[device]
    return ctx


def ptr(a: np.ndarray | None):
    return None if a is None else _vp(a.ctypes.data)
