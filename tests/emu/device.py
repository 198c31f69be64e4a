"""TEST INFRASTRUCTURE ONLY.  Points the product's HOST code (simulation_b200.fd1d / fd2d: ctypes structs, launch
plans, source and phase tables, ping-pong bookkeeping) at the emulated library of tests/emu/build_emu.py, with CPU
torch tensors standing in for device memory, for the duration of one test.  Everything is undone by monkeypatch; the
product itself keeps refusing to run without CUDA (tests/test_capi_symbols.py checks that)."""
from __future__ import annotations

import contextlib
import ctypes as C

from tests.emu import build_emu

_handle = None


def handle():
    global _handle
    if _handle is None:
        from simulation_b200 import _lib
        h = C.CDLL(build_emu.build_library())
        for name, (res, args) in _lib.SYMBOLS.items():
            fn = getattr(h, name)           # every symbol of include/fdtd_b200.h must exist in the emulated build too
            fn.restype, fn.argtypes = res, args
        h.emu_launches.restype = C.c_longlong
        _handle = h
    return _handle


class _Stream:
    """stand-in for torch.cuda.Stream / Event: the emulated library runs every launch at once, in issue order"""
    cuda_stream = 0

    def __init__(self, *a, **k):
        pass

    def wait_stream(self, other):
        pass

    def wait_event(self, ev):
        pass

    def record(self, stream=None):
        pass

    def synchronize(self):
        pass


def install(monkeypatch):
    """-> the emulated library handle; Fdtd1D / Fdtd2D built with device="cpu" now run on it"""
    import torch
    from simulation_b200 import _lib, fd1d, fd2d
    h = handle()
    monkeypatch.setattr(_lib, "_lib", h)
    monkeypatch.setattr(torch.cuda, "is_available", lambda: True)
    monkeypatch.setattr(torch.cuda, "current_device", lambda: 0)
    monkeypatch.setattr(torch.cuda, "device", contextlib.nullcontext)
    monkeypatch.setattr(torch.cuda, "synchronize", lambda *a, **k: None)
    monkeypatch.setattr(torch.cuda, "Stream", _Stream)
    monkeypatch.setattr(torch.cuda, "Event", _Stream)
    monkeypatch.setattr(torch.cuda, "current_stream", lambda *a, **k: _Stream())
    monkeypatch.setattr(torch.cuda, "stream", contextlib.nullcontext)
    no_stream = lambda: C.c_void_p(None)
    accept = lambda *tensors: None
    for mod in (fd1d, fd2d):
        monkeypatch.setattr(mod, "_stream", no_stream)
        monkeypatch.setattr(mod, "_require_cuda", accept)
    h.fdtd2d_tune(0, 0, 0, 0, 0)
    h.fdtd2d_tune2(0, 1)            # deep passes on (default), default halo wait
    h.fdtd2d_tune2(1, 0)
    return h
