"""Training-side feature producer on the GPU: what libDF's ``FftDataset::get_sample``
(libDF/src/dataset.rs:863-914) computes per sample on a CPU worker -- STFT of the clean and the noisy signal, ERB dB
features with exponential mean normalisation, unit-normalised complex features of the first ``nb_spec`` bins -- as
batched device passes over the same kernels the enhancement path uses (``dfb_analysis`` / ``dfb_features``).  Inputs
and outputs are CUDA tensors, so a ``libdfdata``-style loader can hand the model device-resident batches; the on-disk
formats (HDF5) and the augmentation pipeline of the reference stay out of scope.
"""
from __future__ import annotations

from typing import Dict, Optional

import torch
from torch import Tensor

from . import _lib
from ._lib import check
from .libdf import DF


@torch.no_grad()
def fft_features(df: DF, noisy: Tensor, speech: Optional[Tensor] = None, nb_spec: int = 96,
                 norm_alpha: Optional[float] = None) -> Dict[str, Tensor]:
    """noisy / speech float32 CUDA [B, T] -> dict of CUDA tensors in the reference's layouts:
    ``noisy`` c64-as-real [B, 1, Tf, F, 2], ``feat_erb`` [B, 1, Tf, E], ``feat_spec`` [B, 1, Tf, nb_spec, 2]
    (and ``speech`` [B, 1, Tf, F, 2] when given).  Every stream starts from the reset state (dataset.rs:873-876 builds a
    fresh DFState per sample); ``norm_alpha`` defaults to the DF state's (df/utils.py:108-124)."""
    if not noisy.is_cuda or noisy.dtype != torch.float32 or noisy.dim() != 2 or not noisy.is_contiguous():
        raise ValueError("noisy must be a contiguous float32 CUDA tensor [B, T]")
    if noisy.device.index != df.device:
        raise ValueError("noisy lives on another device than the DF state")
    if norm_alpha is None:
        norm_alpha = getattr(df, "norm_alpha", None)
    if norm_alpha is None:
        from .config import ModelConfig
        norm_alpha = ModelConfig(sr=df.sr(), hop_size=df.hop_size()).norm_alpha
    b, t = noisy.shape
    tf, f, e = t // df.hop_size(), df.fft_size() // 2 + 1, df.nb_erb()
    dev = noisy.device
    spec = torch.empty((b, 1, tf, f, 2), dtype=torch.float32, device=dev)
    feat_erb = torch.empty((b, 1, tf, e), dtype=torch.float32, device=dev)
    feat_spec = torch.empty((b, 1, tf, nb_spec, 2), dtype=torch.float32, device=dev)
    L = _lib.lib()
    with torch.cuda.device(dev):
        stream = torch.cuda.current_stream(dev).cuda_stream
        check(L.dfb_features(df.handle, noisy.data_ptr(), b, t, int(nb_spec), float(norm_alpha), spec.data_ptr(),
                             feat_erb.data_ptr(), feat_spec.data_ptr(), stream))
        out = {"noisy": spec, "feat_erb": feat_erb, "feat_spec": feat_spec}
        if speech is not None:
            if speech.shape != noisy.shape or not speech.is_cuda or speech.dtype != torch.float32:
                raise ValueError("speech must match noisy")
            sp = torch.empty_like(spec)
            check(L.dfb_analysis(df.handle, speech.contiguous().data_ptr(), b, t, sp.data_ptr(), stream))
            out["speech"] = sp
    return out
