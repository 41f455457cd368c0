"""ctypes binding of libqmcb.so (include/qmcb.h).  No CPU fallback: if the CUDA library is
missing or a call fails, a RuntimeError is raised."""
import ctypes as C
import os

import numpy as np
import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("QMCB_LIB") or os.path.join(_HERE, "lib", "libqmcb.so")

RADIAL = {"gto_pure": 0, "gto": 1, "sto_pure": 2, "sto": 3}


class QmcbSystem(C.Structure):
    _fields_ = [
        ("nelec", C.c_int32), ("nup", C.c_int32), ("ndown", C.c_int32), ("natom", C.c_int32),
        ("nbas", C.c_int32), ("nao", C.c_int32), ("nmo", C.c_int32), ("radial_type", C.c_int32),
        ("contract", C.c_int32),
        ("atom_coords", C.c_void_p), ("atomic_number", C.c_void_p), ("bas_atom", C.c_void_p),
        ("bas_exp", C.c_void_p), ("bas_coeffs", C.c_void_p), ("bas_norm", C.c_void_p),
        ("bas_kx", C.c_void_p), ("bas_ky", C.c_void_p), ("bas_kz", C.c_void_p), ("bas_kr", C.c_void_p),
        ("index_ctr", C.c_void_p), ("mo", C.c_void_p),
        ("nconf", C.c_int32), ("cfg_up", C.c_void_p), ("cfg_down", C.c_void_p), ("ci", C.c_void_p),
        ("use_jee", C.c_int32), ("jee_w", C.c_double), ("use_jen", C.c_int32), ("jen_w", C.c_double),
        ("gram_fma", C.c_int32),
        ("een_nterm", C.c_int32), ("een_num", C.c_void_p), ("een_denom", C.c_void_p), ("een_fc", C.c_void_p),
    ]


_lib = None

_SIGS = {
    "qmcb_abi_version": (C.c_int, []),
    "qmcb_last_error": (C.c_char_p, []),
    "qmcb_plan_create": (C.c_int, [C.POINTER(QmcbSystem), C.c_int, C.POINTER(C.c_void_p)]),
    "qmcb_plan_update": (C.c_int, [C.c_void_p, C.POINTER(QmcbSystem)]),
    "qmcb_plan_destroy": (None, [C.c_void_p]),
    "qmcb_plan_info": (C.c_int, [C.c_void_p, C.c_int]),
    "qmcb_psi": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]),
    "qmcb_local_energy": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p,
                                    C.c_void_p, C.c_void_p]),
    "qmcb_grad_psi": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_void_p, C.c_void_p]),
    "qmcb_metropolis_step": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p,
                                       C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_double,
                                       C.c_double, C.c_uint64, C.c_uint64, C.c_void_p, C.c_void_p,
                                       C.c_void_p]),
    "qmcb_backward_workspace_bytes": (C.c_int64, [C.c_void_p, C.c_int64]),
    "qmcb_psi_backward": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64] + [C.c_void_p] * 9),
    "qmcb_local_energy_backward_workspace_bytes": (C.c_int64, [C.c_void_p, C.c_int64]),
    "qmcb_local_energy_backward": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64] +
                                   [C.c_void_p] * 9),
    "qmcb_stats_workspace_bytes": (C.c_int64, [C.c_int64]),
    "qmcb_energy_stats": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]),
    "qmcb_local_energy_stats": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p,
                                          C.c_void_p, C.c_void_p, C.c_void_p]),
    "qmcb_ao": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_void_p, C.c_void_p,
                          C.c_void_p, C.c_void_p]),
    "qmcb_mo": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]),
    "qmcb_jastrow": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_void_p, C.c_void_p,
                               C.c_void_p, C.c_void_p]),
    "qmcb_slater": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_void_p,
                              C.c_void_p, C.c_void_p]),
    "qmcb_fp64_probe": (C.c_int, [C.c_int, C.c_int64, C.c_void_p, C.POINTER(C.c_double), C.c_void_p]),
}

EXPORTS = tuple(_SIGS)


def lib():
    """Loads libqmcb.so once.  Raises if it has not been built (python -m qmctorch_b200.build)."""
    global _lib
    if _lib is None:
        if not os.path.isfile(LIB_PATH):
            raise RuntimeError(
                "qmctorch_b200: %s is missing - build it with `python -m qmctorch_b200.build` "
                "(there is no CPU fallback)" % LIB_PATH)
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in _SIGS.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        if L.qmcb_abi_version() != 1:
            raise RuntimeError("qmctorch_b200: libqmcb.so ABI version mismatch")
        _lib = L
    return _lib


def check(rc, what):
    if rc != 0:
        msg = lib().qmcb_last_error()
        raise RuntimeError("%s failed (code %d): %s" % (what, rc, (msg or b"").decode() or "CUDA error"))


def ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def stream_ptr(device):
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def _np(a, dt):
    return np.ascontiguousarray(np.asarray(a), dtype=dt)


class SystemArrays:
    """Keeps the host arrays of a qmcb_system alive and exposes the ctypes struct."""

    def __init__(self, **kw):
        self.keep = {}
        s = QmcbSystem()
        for k in ("nelec", "nup", "ndown", "natom", "nbas", "nao", "nmo", "radial_type", "contract",
                  "nconf", "use_jee", "use_jen", "gram_fma"):
            setattr(s, k, int(kw[k]))
        s.jee_w = float(kw["jee_w"])
        s.jen_w = float(kw["jen_w"])
        s.een_nterm = int(kw.get("een_nterm", 0))
        for k in ("een_num", "een_denom", "een_fc"):
            a = _np(kw.get(k, np.zeros(1)), np.float64)
            self.keep[k] = a
            setattr(s, k, a.ctypes.data)
        for k in ("atom_coords", "atomic_number", "bas_exp", "bas_coeffs", "bas_norm", "mo", "ci"):
            a = _np(kw[k], np.float64)
            self.keep[k] = a
            setattr(s, k, a.ctypes.data)
        for k in ("bas_atom", "bas_kx", "bas_ky", "bas_kz", "bas_kr", "index_ctr", "cfg_up", "cfg_down"):
            a = _np(kw[k], np.int32)
            self.keep[k] = a
            setattr(s, k, a.ctypes.data)
        self.struct = s
