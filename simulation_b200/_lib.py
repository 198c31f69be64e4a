"""ctypes binding of ``csrc/libfdtd_b200.so`` (C ABI declared in ``include/fdtd_b200.h``).

There is NO fallback: if the shared library is missing or a call fails, this raises.  Build it with
``make -C simulation_b200/csrc`` (or ``python -c "import __graft_entry__ as g; g.build()"``).
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "csrc", "libfdtd_b200.so")

F32, F64 = 0, 1
TFSF, LOSSY, ABC, FLUX, DEBYE, LAZY_EZ, GHOST_DECAY, INCIDENT_READY = 1, 2, 4, 8, 16, 32, 64, 128
DZ, EZ, HX, HY, IHX, IHY, IZ, NFIELDS = range(8)


class FdtdError(RuntimeError):
    pass


class PmlLayer(C.Structure):          # fdtd_pmlayer: reference declaration order
    _fields_ = [(n, C.c_void_p) for n in ("fx1", "fx2", "fx3", "fy1", "fy2", "fy3", "gx2", "gx3", "gy2", "gy3")]


class Medium2D(C.Structure):
    _fields_ = [("naz", C.c_void_p), ("nbz", C.c_void_p)]


class Medium1D(C.Structure):
    _fields_ = [("nax", C.c_void_p), ("nbx", C.c_void_p), ("ncx", C.c_void_p), ("ndx", C.c_void_p)]


class Source(C.Structure):
    _fields_ = [("target", C.c_void_p), ("index", C.c_longlong), ("hard", C.c_int), ("value", C.c_double)]


class FTrans(C.Structure):
    _fields_ = [("r_pt", C.c_void_p), ("i_pt", C.c_void_p), ("r_in", C.c_void_p), ("i_in", C.c_void_p)]


class Problem1D(C.Structure):
    _fields_ = [("dtype", C.c_int), ("nx", C.c_int), ("flags", C.c_int),
                ("ca", C.c_void_p), ("cb", C.c_void_p), ("md", Medium1D),
                ("state", (C.c_void_p * 5) * 2), ("bc", C.c_void_p * 2),
                ("src_field", C.c_int), ("src_index", C.c_int), ("src_hard", C.c_int),
                ("nf", C.c_int), ("dft_sample", C.c_int), ("ft", FTrans),
                ("dft_cos", C.POINTER(C.c_double)), ("dft_sin", C.POINTER(C.c_double))]


class Problem2D(C.Structure):
    _fields_ = [("dtype", C.c_int), ("nx", C.c_int), ("ny", C.c_int),
                ("row_lo", C.c_int), ("row_hi", C.c_int), ("row_base", C.c_int), ("rows_alloc", C.c_int),
                ("npml", C.c_int), ("flags", C.c_int),
                ("pml", PmlLayer), ("md", Medium2D),
                ("state", (C.c_void_p * NFIELDS) * 2),
                ("ezi", C.c_void_p), ("hxi", C.c_void_p), ("bc", C.c_void_p),
                ("ezi_hist", C.c_void_p), ("hxi_hist", C.c_void_p),
                ("src_i", C.c_int), ("src_j", C.c_int), ("src_hard", C.c_int),
                ("ident_row_lo", C.c_int), ("ident_row_hi", C.c_int), ("ident_col_lo", C.c_int), ("ident_col_hi", C.c_int),
                ("nf", C.c_int), ("ft", FTrans), ("dft_cos", C.POINTER(C.c_double)), ("dft_sin", C.POINTER(C.c_double)),
                ("halo", C.c_int), ("peer_up", (C.c_void_p * NFIELDS) * 2), ("peer_dn", (C.c_void_p * NFIELDS) * 2),
                ("peer_up_base", C.c_int), ("peer_dn_base", C.c_int),
                ("sync_local", C.c_void_p), ("sync_up", C.c_void_p), ("sync_dn", C.c_void_p),
                ("epoch", C.c_ulonglong),
                ("lossy_row_lo", C.c_int), ("lossy_row_hi", C.c_int), ("lossy_col_lo", C.c_int), ("lossy_col_hi", C.c_int)]


# every symbol include/fdtd_b200.h declares: name -> (restype, argtypes)
_P, _I, _D = C.c_void_p, C.c_int, C.c_double
SYMBOLS = {
    "fdtd_last_error": (C.c_char_p, []),
    "fdtd_version": (_I, []),
    "fdtd_device_info": (_I, [C.POINTER(_I), C.POINTER(C.c_size_t), C.POINTER(C.c_size_t)]),
    "fdtd_malloc": (_I, [C.POINTER(_P), C.c_size_t]),
    "fdtd_free": (_I, [_P]),
    "fdtd_memset0": (_I, [_P, C.c_size_t, _P]),
    "fdtd_upload": (_I, [_P, _P, C.c_size_t, _P]),
    "fdtd_download": (_I, [_P, _P, C.c_size_t, _P]),
    "fdtd_stream_sync": (_I, [_P]),
    "fdtd_enable_peer_access": (_I, [_I]),
    "fdtd_ipc_export": (_I, [_P, _P]),
    "fdtd_ipc_open": (_I, [_P, C.POINTER(_P)]),
    "fdtd_ipc_close": (_I, [_P]),
    "fdtd1d_exfield": (_I, [_I, _I, _P, _P, _P, _P, C.POINTER(Source), _P]),
    "fdtd1d_hyfield": (_I, [_I, _I, _P, _P, _P, _I, _P]),
    "fdtd1d_dxfield": (_I, [_I, _I, _P, _P, C.POINTER(Source), _P]),
    "fdtd1d_exfield_flux": (_I, [_I, _I, C.POINTER(Medium1D), _P, _P, _P, _P, _P]),
    "fdtd1d_fourier": (_I, [_I, _I, _I, C.POINTER(_D), C.POINTER(_D), _P, _I, C.POINTER(FTrans), _P]),
    "fdtd2d_fourier": (_I, [_I, _I, _I, _I, C.POINTER(_D), C.POINTER(_D), _P, _I, _P, C.POINTER(FTrans), _P]),
    "fdtd1d_advance": (_I, [C.POINTER(Problem1D), _I, _I, C.POINTER(_D), _I, _P, C.POINTER(_I)]),
    "fdtd2d_ezinct": (_I, [_I, _I, _P, _P, _P, _P]),
    "fdtd2d_dfield": (_I, [_I, _I, _I, C.POINTER(PmlLayer), _P, _P, _P, C.POINTER(Source), _P]),
    "fdtd2d_inctdz": (_I, [_I, _I, _I, _I, _P, _P, _P]),
    "fdtd2d_efield": (_I, [_I, _I, _I, C.POINTER(Medium2D), _P, _P, _P, _P]),
    "fdtd2d_hxinct": (_I, [_I, _I, _P, _P, _P]),
    "fdtd2d_hfield": (_I, [_I, _I, _I, C.POINTER(PmlLayer), _P, _P, _P, _P, _P, _P]),
    "fdtd2d_incthx": (_I, [_I, _I, _I, _I, _P, _P, _P]),
    "fdtd2d_incthy": (_I, [_I, _I, _I, _I, _P, _P, _P]),
    "fdtd2d_dielectric_cylinder": (_I, [_I, _I, _I, _I, _I, _D, _D, _D, _I, _I, _P, _P, _P]),
    "fdtd2d_pmlparam": (_I, [_I, _I, _I, _I, _P, C.POINTER(PmlLayer), _P]),
    "fdtd2d_advance": (_I, [C.POINTER(Problem2D), _I, _I, C.POINTER(_D), _I, _P, C.POINTER(_I)]),
    "fdtd2d_check_identity": (_I, [C.POINTER(Problem2D), C.POINTER(C.c_longlong)]),
    "fdtd2d_check_lossless_outside": (_I, [C.POINTER(Problem2D), C.POINTER(C.c_longlong)]),
    "fdtd2d_preload": (_I, [_I, _I, _I]),
    "fdtd2d_max_tblock": (_I, [_I, _I]),
    "fdtd2d_tune": (_I, [_I, _I, _I, _I, _I]),
    "fdtd2d_tune2": (_I, [_I, C.c_longlong]),
    "fdtd2d_plan": (_I, [C.POINTER(Problem2D), _I, _I, C.POINTER(_I), _I, C.POINTER(_I), C.POINTER(_I)]),
    "fdtd2d_incident_line": (_I, [C.POINTER(Problem2D), _I, C.POINTER(_D), _P, _P, _P]),
    "fdtd2d_halo_status": (_I, [C.POINTER(Problem2D), C.POINTER(C.c_ulonglong)]),
}
TUNE_DEEP, TUNE_HALO_WAIT_MS, TUNE_VARIANT, TUNE_EDGE_CHUNKS, TUNE_COL_FAST = 0, 1, 2, 3, 4

_lib = None


def lib():
    """Load (once) and return the shared library with prototypes installed.  Raises if it is missing."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise FdtdError(f"{LIB_PATH} is missing: the CUDA extension has not been built "
                            "(make -C simulation_b200/csrc).  There is no CPU fallback.")
        h = C.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(h, name)          # AttributeError if the .so does not export a declared symbol
            fn.restype, fn.argtypes = res, args
        _lib = h
    return _lib


def check(rc: int, what: str = ""):
    if rc != 0:
        msg = lib().fdtd_last_error().decode(errors="replace")
        raise FdtdError(f"{what or 'libfdtd_b200'} failed (code {rc}): {msg}")


def dtype_code(np_dtype) -> int:
    import numpy as np
    dt = np.dtype(np_dtype)
    if dt == np.float32:
        return F32
    if dt == np.float64:
        return F64
    raise FdtdError(f"unsupported dtype {dt}: the path computes in float32 or float64")


class DeviceBuffer:
    """Device memory from the library's own allocator (cudaMalloc), exportable over CUDA IPC; torch wraps it without
    a copy through ``__cuda_array_interface__``.  Freed with the object."""

    def __init__(self, shape, np_dtype):
        import numpy as np
        self.shape = tuple(int(x) for x in shape)
        self.np_dtype = np.dtype(np_dtype)
        self.nbytes = int(np.prod(self.shape)) * self.np_dtype.itemsize
        p = C.c_void_p()
        check(lib().fdtd_malloc(C.byref(p), max(self.nbytes, 16)), "fdtd_malloc")
        self.ptr = p.value
        check(lib().fdtd_memset0(C.c_void_p(self.ptr), max(self.nbytes, 16), None), "fdtd_memset0")
        check(lib().fdtd_stream_sync(None), "fdtd_stream_sync")

    @property
    def __cuda_array_interface__(self):
        return {"shape": self.shape, "typestr": self.np_dtype.str, "data": (self.ptr, False), "version": 3, "strides": None}

    def ipc_handle(self) -> bytes:
        buf = C.create_string_buffer(64)
        check(lib().fdtd_ipc_export(C.c_void_p(self.ptr), buf), "fdtd_ipc_export")
        return buf.raw

    def __del__(self):
        try:
            if getattr(self, "ptr", None):
                lib().fdtd_free(C.c_void_p(self.ptr))
                self.ptr = None
        except Exception:
            pass


def ipc_close(mapped: int) -> None:
    check(lib().fdtd_ipc_close(C.c_void_p(mapped)), "fdtd_ipc_close")


def ipc_open(handle: bytes) -> int:
    out = C.c_void_p()
    check(lib().fdtd_ipc_open(C.create_string_buffer(handle, 64), C.byref(out)), "fdtd_ipc_open")
    return int(out.value)
