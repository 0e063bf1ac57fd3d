"""ctypes binding of the CPU oracle (oracle/libdf_oracle.c) shaped like the reference's
python module ``libdf`` (pyDF/src/lib.rs:14-307, stubs pyDF/libdf.pyi).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's CPU
baseline leg.  The product package (deepfilternet_b200) never imports this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import Optional

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libdfo.so")


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "libdf_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "libdfo.so"], stdout=subprocess.DEVNULL)
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        fp = C.POINTER(C.c_float)
        ip = C.POINTER(C.c_int64)
        L.dfo_create.restype = C.c_void_p
        L.dfo_create.argtypes = [C.c_int] * 5
        L.dfo_free.argtypes = [C.c_void_p]
        L.dfo_reset.argtypes = [C.c_void_p]
        L.dfo_get_erb_widths.argtypes = [C.c_void_p, ip]
        L.dfo_get_window.argtypes = [C.c_void_p, fp]
        L.dfo_get_wnorm.argtypes = [C.c_void_p]
        L.dfo_get_wnorm.restype = C.c_float
        L.dfo_erb_widths.argtypes = [C.c_int] * 4 + [ip]
        L.dfo_analysis.argtypes = [C.c_void_p, fp, C.c_int64, C.c_int64, C.c_int, fp]
        L.dfo_synthesis.argtypes = [C.c_void_p, fp, C.c_int64, C.c_int64, C.c_int, fp]
        L.dfo_erb.argtypes = [fp, C.c_int64, C.c_int64, ip, C.c_int, C.c_int, fp]
        L.dfo_erb.restype = C.c_int
        L.dfo_erb_inv.argtypes = [fp, C.c_int64, ip, C.c_int, fp]
        L.dfo_apply_erb_gains.argtypes = [fp, fp, C.c_int64, ip, C.c_int]
        L.dfo_erb_norm.argtypes = [fp, C.c_int64, C.c_int64, C.c_int64, C.c_float, fp]
        L.dfo_unit_norm.argtypes = [fp, C.c_int64, C.c_int64, C.c_int64, C.c_float, fp]
        L.dfo_unit_norm_init.argtypes = [C.c_int64, fp]
        L.dfo_mean_norm_init.argtypes = [C.c_int64, fp]
        _lib = L
    return _lib


def _fp(a: np.ndarray):
    return a.ctypes.data_as(C.POINTER(C.c_float))


def _ip(a: np.ndarray):
    return a.ctypes.data_as(C.POINTER(C.c_int64))


def _require_contig(a: np.ndarray):
    if a.size == 0 or not a.flags["C_CONTIGUOUS"]:
        # pyDF/src/lib.rs:59-64
        raise RuntimeError("[df] Input array empty or not contiguous.")


class DF:
    """Mirror of pyDF ``DF`` (pyDF/src/lib.rs:14-136)."""

    def __init__(self, sr: int, fft_size: int, hop_size: int, nb_bands: Optional[int] = 32,
                 min_nb_erb_freqs: Optional[int] = 1):
        nb_bands = 32 if nb_bands is None else nb_bands
        min_nb_erb_freqs = 1 if min_nb_erb_freqs is None else min_nb_erb_freqs
        self._h = lib().dfo_create(sr, fft_size, hop_size, nb_bands, min_nb_erb_freqs)
        if not self._h:
            raise RuntimeError("assertion failed: hop_size * 2 <= fft_size")
        self._sr, self._fft, self._hop, self._nb = sr, fft_size, hop_size, nb_bands

    def __del__(self):
        if getattr(self, "_h", None):
            lib().dfo_free(self._h)
            self._h = None

    def analysis(self, input: np.ndarray, reset: Optional[bool] = True) -> np.ndarray:
        x = np.asarray(input)
        if x.dtype != np.float32 or x.ndim != 2:
            raise TypeError("analysis expects float32 [C, T]")
        _require_contig(x)
        c, t = x.shape
        tf = t // self._hop
        out = np.zeros((c, tf, self._fft // 2 + 1), dtype=np.complex64)
        lib().dfo_analysis(self._h, _fp(x), c, t, 1 if (reset is None or reset) else 0,
                           _fp(out.view(np.float32)))
        return out

    def synthesis(self, input: np.ndarray, reset: Optional[bool] = True) -> np.ndarray:
        x = np.asarray(input)
        if x.dtype != np.complex64 or x.ndim != 3:
            raise TypeError("synthesis expects complex64 [C, T, F]")
        _require_contig(x)
        c, tf, f = x.shape
        assert f == self._fft // 2 + 1
        out = np.zeros((c, tf * self._hop), dtype=np.float32)
        lib().dfo_synthesis(self._h, _fp(x.view(np.float32)), c, tf,
                            1 if (reset is None or reset) else 0, _fp(out))
        return out

    def erb_widths(self) -> np.ndarray:
        out = np.zeros(self._nb, dtype=np.int64)
        lib().dfo_get_erb_widths(self._h, _ip(out))
        return out.astype(np.uint64)

    def fft_window(self) -> np.ndarray:
        out = np.zeros(self._fft, dtype=np.float32)
        lib().dfo_get_window(self._h, _fp(out))
        return out

    def sr(self) -> int:
        return self._sr

    def fft_size(self) -> int:
        return self._fft

    def hop_size(self) -> int:
        return self._hop

    def nb_erb(self) -> int:
        return self._nb

    def reset(self) -> None:
        lib().dfo_reset(self._h)


def erb_widths(sr: int, fft_size: int, nb_bands: int, min_nb_freqs: int) -> np.ndarray:
    out = np.zeros(nb_bands, dtype=np.int64)
    lib().dfo_erb_widths(sr, fft_size, nb_bands, min_nb_freqs, _ip(out))
    return out


def erb(input: np.ndarray, erb_fb: np.ndarray, db: Optional[bool] = True) -> np.ndarray:
    """pyDF/src/lib.rs:142-192"""
    x = np.ascontiguousarray(input, dtype=np.complex64)
    if x.ndim not in (2, 3, 4):
        raise ValueError(f"Dimension not supported for erb: {x.ndim}")
    fb = np.ascontiguousarray(erb_fb).astype(np.int64)
    f = x.shape[-1]
    n = x.size // f
    out = np.zeros(x.shape[:-1] + (len(fb),), dtype=np.float32)
    rc = lib().dfo_erb(_fp(x.view(np.float32)), n, f, _ip(fb), len(fb),
                       1 if (db is None or db) else 0, _fp(out))
    if rc != 0:
        raise RuntimeError("DF shape error: erb widths do not sum to the number of frequency bins")
    return out


def erb_inv(input: np.ndarray, erb_fb: np.ndarray) -> np.ndarray:
    """pyDF/src/lib.rs:194-250"""
    x = np.ascontiguousarray(input, dtype=np.float32)
    fb = np.ascontiguousarray(erb_fb).astype(np.int64)
    if x.shape[-1] != len(fb):
        raise ValueError(
            f"Number of erb bands do not match with input: {x.shape[-1]}, {len(fb)}")
    if x.ndim not in (2, 3, 4):
        raise ValueError(f"Dimension not supported for erb: {x.ndim}")
    n = x.size // len(fb)
    out = np.zeros(x.shape[:-1] + (int(fb.sum()),), dtype=np.float32)
    lib().dfo_erb_inv(_fp(x), n, _ip(fb), len(fb), _fp(out))
    return out


def erb_norm(erb: np.ndarray, alpha: float, state: Optional[np.ndarray] = None) -> np.ndarray:
    """pyDF/src/lib.rs:252-274 (the reference also clobbers its argument; the copy returned is
    what callers use)."""
    x = np.array(erb, dtype=np.float32, order="C", copy=True)
    assert x.ndim == 3
    st = None if state is None else np.ascontiguousarray(state, dtype=np.float32)
    lib().dfo_erb_norm(_fp(x), x.shape[0], x.shape[1], x.shape[2], float(alpha),
                       _fp(st) if st is not None else None)
    return x


def unit_norm(spec: np.ndarray, alpha: float, state: Optional[np.ndarray] = None) -> np.ndarray:
    """pyDF/src/lib.rs:276-298"""
    x = np.array(spec, dtype=np.complex64, order="C", copy=True)
    assert x.ndim == 3
    st = None if state is None else np.ascontiguousarray(state, dtype=np.float32)
    lib().dfo_unit_norm(_fp(x.view(np.float32)), x.shape[0], x.shape[1], x.shape[2], float(alpha),
                        _fp(st) if st is not None else None)
    return x


def unit_norm_init(num_freq_bins: int) -> np.ndarray:
    """pyDF/src/lib.rs:300-307"""
    out = np.zeros((1, num_freq_bins), dtype=np.float32)
    lib().dfo_unit_norm_init(num_freq_bins, _fp(out))
    return out
