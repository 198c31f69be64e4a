"""fdtd-b200: B200-native (sm_100a) kernels for the Yee time-stepping hot path of dsarvan/simulation.

``surface``  setup surface kept from the reference (waveforms, PML vectors, media) -- host numpy
``fd1d``     1D Ex/Hy line: reference-named step functions + fused time-blocked ``Fdtd1D.advance``
``fd2d``     2D TM grid:   reference-named step functions + fused time-blocked ``Fdtd2D.advance``
``slab``     row-slab decomposition of a 2D grid over the GPUs of one node (torch.distributed)
``_lib``     ctypes binding of csrc/libfdtd_b200.so (C ABI: include/fdtd_b200.h); no CPU fallback

Importing the package does not import torch; ``fd1d`` / ``fd2d`` / ``slab`` do (device memory, streams).
"""
from . import _lib, surface  # noqa: F401

__all__ = ["_lib", "surface", "fd1d", "fd2d", "slab"]
__version__ = "0.1.0"
