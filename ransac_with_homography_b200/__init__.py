"""ransac_with_homography_b200 -- B200 (sm_100a) implementation of the geometric hot path of
choice17/ransac_with_homography: RANSAC homography estimation and backward perspective warp.

``homography`` and ``ransac`` mirror the reference's modules of the same names; ``device`` is the
device-resident / batched / multi-GPU interface used for throughput work.  All arithmetic runs in
hand-written CUDA kernels behind the C ABI of ``libb200geo.so`` (``include/b200geo.h``); there is
no CPU fallback.
"""
from . import _cabi  # noqa: F401

__version__ = "0.1.0"
__all__ = ["homography", "ransac", "device", "_cabi"]


def __getattr__(name):
    if name in ("homography", "ransac", "device"):
        import importlib
        return importlib.import_module(f"{__name__}.{name}")
    raise AttributeError(name)
