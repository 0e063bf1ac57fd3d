"""Frame-incremental enhancement with carried state: host-side mirror of the reference's streaming runtime
(``DfTract::process`` libDF/src/tract.rs:509-642; C ABI libDF/src/capi.rs:83-253) for B independent streams at once.

    s = DfStream(model, df_state, batch=4)
    for chunk in chunks:                 # chunk: float32 [4, n * hop], any n >= 1
        out = s.process(chunk)           # [4, n * hop], trailing the input by s.latency_frames hops
    tail = s.flush()                     # [4, latency * hop]

The concatenation of the outputs equals ``enhance(model, df_state, audio, pad=False)`` delayed by
``latency_frames * hop`` samples.  Everything runs in the CUDA library (dfb_stream_* in include/dfb200.h).
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch
from torch import Tensor

from . import _lib
from ._lib import check
from .libdf import DF
from .model import DfNet


class DfStream:
    def __init__(self, model: DfNet, df_state: DF, batch: int = 1, atten_lim_db: Optional[float] = None):
        self.model, self.df_state, self.batch = model, df_state, int(batch)
        h = C.c_void_p()
        lim = abs(float(atten_lim_db)) if atten_lim_db is not None else 0.0
        check(_lib.lib().dfb_stream_create(C.byref(h), model.handle, df_state.handle, self.batch, lim))
        self._h = h
        self.hop = int(_lib.lib().dfb_stream_frame_length(h))
        self.latency_frames = int(_lib.lib().dfb_stream_latency_frames(h))

    def __del__(self):
        h = getattr(self, "_h", None)
        if h:
            try:
                _lib.lib().dfb_stream_free(h)
            except Exception:
                pass
            self._h = None

    def set_lsnr_thresholds(self, min_db_thresh: float = -10.0, max_db_erb_thresh: float = 30.0,
                            max_db_df_thresh: float = 20.0, enable: bool = True) -> None:
        """Stage gating of the Rust runtime (tract.rs:658-672, defaults tract.rs:180-185); off unless called."""
        check(_lib.lib().dfb_stream_set_lsnr_thresholds(self._h, int(enable), float(min_db_thresh), float(max_db_erb_thresh),
                                                        float(max_db_df_thresh)))

    def reset(self) -> None:
        check(_lib.lib().dfb_stream_reset(self._h))

    @torch.no_grad()
    def process(self, audio: Tensor) -> Tensor:
        """audio float32 [B, n * hop] (CPU or the model's CUDA device) -> enhanced [B, n * hop] on the same device."""
        if audio.dim() != 2 or audio.shape[0] != self.batch or audio.shape[1] == 0 or audio.shape[1] % self.hop:
            raise ValueError(f"audio must have shape [{self.batch}, n * {self.hop}]")
        n = audio.shape[1] // self.hop
        if audio.is_cuda:
            if audio.device != self.model.cuda_device:
                raise ValueError("audio lives on another device than the model")
            x = audio.to(torch.float32).contiguous()
            out = torch.empty_like(x)
            with torch.cuda.device(x.device):
                check(_lib.lib().dfb_stream_process(self._h, x.data_ptr(), n, out.data_ptr(),
                                                    torch.cuda.current_stream(x.device).cuda_stream))
            return out
        x = audio.detach().to("cpu", torch.float32).contiguous()
        out = torch.empty_like(x)
        check(_lib.lib().dfb_stream_process_host(self._h, x.data_ptr(), n, out.data_ptr()))
        return out

    @torch.no_grad()
    def flush(self) -> Tensor:
        """The ``latency_frames`` hops still in flight at the end of the stream (CPU tensor)."""
        out = torch.zeros((self.batch, self.latency_frames * self.hop), dtype=torch.float32)
        if self.latency_frames:
            check(_lib.lib().dfb_stream_process_host(self._h, None, 0, out.data_ptr()))
        return out
