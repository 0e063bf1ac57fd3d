"""Drop-in for ``df.io`` (DeepFilterNet/df/io.py:25-129): ``load_audio``, ``save_audio``, ``resample``,
``get_resample_params`` with the reference's signatures.

* File I/O: the reference goes through ``torchaudio.load / save`` (which need a codec backend that is not in this
  image); WAV files (PCM 8/16/24/32 bit, IEEE float32) are read and written here with a small RIFF parser, with
  torchaudio's normalisation (int16 / 32768 ...).
* Resampling: the taps are ``torchaudio.functional.resample``'s (``_get_sinc_resample_kernel``, restated below: hann /
  kaiser windowed sinc for the gcd-reduced rates, computed on the host in float64 like torchaudio does) and the
  polyphase convolution runs in the CUDA library (``dfb_resample_host`` -> ``k_resample``); no CPU compute path.
"""
from __future__ import annotations

import math
import os
import struct
from collections import namedtuple
from typing import Any, Dict, Optional, Tuple, Union

import numpy as np
import torch
from torch import Tensor

from . import _lib
from ._lib import check

AudioMetaData = namedtuple("AudioMetaData", ["sample_rate", "num_frames", "num_channels", "bits_per_sample", "encoding"])

TA_RESAMPLE_SINC = "sinc_interp_hann"
TA_RESAMPLE_KAISER = "sinc_interp_kaiser"


# ------------------------------------------------------------------------------------------ WAV ----
def _read_wav(path: str):
    with open(path, "rb") as f:
        data = f.read()
    if data[:4] != b"RIFF" or data[8:12] != b"WAVE":
        raise RuntimeError(f"{path}: not a RIFF/WAVE file (only WAV is readable without a torchaudio backend)")
    pos, fmt, pcm = 12, None, None
    while pos + 8 <= len(data):
        cid, size = data[pos:pos + 4], struct.unpack("<I", data[pos + 4:pos + 8])[0]
        body = data[pos + 8:pos + 8 + size]
        if cid == b"fmt ":
            tag, ch, sr, _, _, bits = struct.unpack("<HHIIHH", body[:16])
            if tag == 0xFFFE and len(body) >= 26:  # WAVE_FORMAT_EXTENSIBLE: the sub-format's first two bytes
                tag = struct.unpack("<H", body[24:26])[0]
            fmt = (tag, ch, sr, bits)
        elif cid == b"data":
            pcm = body
        pos += 8 + size + (size & 1)
    if fmt is None or pcm is None:
        raise RuntimeError(f"{path}: missing fmt / data chunk")
    tag, ch, sr, bits = fmt
    if tag == 1:
        if bits == 8:
            x = (np.frombuffer(pcm, dtype=np.uint8).astype(np.float32) - 128.0) / 128.0
        elif bits == 16:
            x = np.frombuffer(pcm, dtype="<i2").astype(np.float32) / 32768.0
        elif bits == 24:
            b = np.frombuffer(pcm[:len(pcm) // 3 * 3], dtype=np.uint8).reshape(-1, 3).astype(np.int32)
            v = b[:, 0] | (b[:, 1] << 8) | (b[:, 2] << 16)
            x = np.where(v >= 1 << 23, v - (1 << 24), v).astype(np.float32) / float(1 << 23)
        elif bits == 32:
            x = np.frombuffer(pcm, dtype="<i4").astype(np.float32) / float(1 << 31)
        else:
            raise RuntimeError(f"{path}: unsupported PCM width {bits}")
        enc = "PCM_S" if bits > 8 else "PCM_U"
    elif tag == 3 and bits == 32:
        x, enc = np.frombuffer(pcm, dtype="<f4").astype(np.float32), "PCM_F"
    else:
        raise RuntimeError(f"{path}: unsupported WAV format tag {tag} / {bits} bit")
    n = len(x) // ch
    return x[:n * ch].reshape(n, ch).T.copy(), AudioMetaData(sr, n, ch, bits, enc)


def _write_wav(path: str, audio: np.ndarray, sr: int) -> None:
    ch, n = audio.shape
    if audio.dtype == np.int16:
        tag, bits, raw = 1, 16, audio.T.astype("<i2").tobytes()
    elif audio.dtype == np.float32:
        tag, bits, raw = 3, 32, audio.T.astype("<f4").tobytes()
    else:
        raise ValueError(f"unsupported sample type {audio.dtype}")
    hdr = struct.pack("<4sI4s4sIHHIIHH4sI", b"RIFF", 36 + len(raw), b"WAVE", b"fmt ", 16, tag, ch, sr, sr * ch * bits // 8,
                      ch * bits // 8, bits, b"data", len(raw))
    with open(path, "wb") as f:
        f.write(hdr + raw)


def load_audio(file: str, sr: Optional[int] = None, verbose=True, **kwargs) -> Tuple[Tensor, AudioMetaData]:
    """df/io.py:25-59: audio float32 [C, T] (resampled to ``sr`` when given) and the ORIGINAL file's meta data.
    ``frame_offset`` / ``num_frames`` / ``method`` keyword arguments as in the reference."""
    rkwargs = {}
    if "method" in kwargs:
        rkwargs["method"] = kwargs.pop("method")
    x, info = _read_wav(file)
    off = int(kwargs.get("frame_offset", 0))
    num = int(kwargs.get("num_frames", -1))
    if num >= 0 and sr is not None:
        num *= info.sample_rate // sr        # io.py:48-49
    x = x[:, off:off + num] if num >= 0 else x[:, off:]
    audio = torch.from_numpy(np.ascontiguousarray(x))
    if sr is not None and info.sample_rate != sr:
        if verbose:
            import warnings
            warnings.warn(f"Audio sampling rate does not match model sampling rate ({info.sample_rate}, {sr}). Resampling...")
        audio = resample(audio, info.sample_rate, sr, **rkwargs)
    return audio.contiguous(), info


def save_audio(file: str, audio: Union[Tensor, np.ndarray], sr: int, output_dir: Optional[str] = None,
               suffix: Optional[str] = None, log: bool = False, dtype=torch.int16):
    """df/io.py:62-86 (int16 scaling by 1 << 15 included)."""
    outpath = file
    if suffix is not None:
        file, ext = os.path.splitext(file)
        outpath = file + f"_{suffix}" + ext
    if output_dir is not None:
        outpath = os.path.join(output_dir, os.path.basename(outpath))
    if log:
        import logging
        logging.getLogger("deepfilternet_b200").info("Saving audio file '%s'", outpath)
    audio = torch.as_tensor(audio)
    if audio.ndim == 1:
        audio = audio.unsqueeze(0)
    if dtype == torch.int16 and audio.dtype != torch.int16:
        audio = (audio * (1 << 15)).to(torch.int16)
    if dtype == torch.float32 and audio.dtype != torch.float32:
        audio = audio.to(torch.float32) / (1 << 15)
    _write_wav(outpath, audio.cpu().numpy(), int(sr))
    return outpath


# ------------------------------------------------------------------------------------- resample ----
def get_resample_params(method: str) -> Dict[str, Any]:
    """df/io.py:95-113."""
    params = {
        "sinc_fast": {"resampling_method": TA_RESAMPLE_SINC, "lowpass_filter_width": 16},
        "sinc_best": {"resampling_method": TA_RESAMPLE_SINC, "lowpass_filter_width": 64},
        "kaiser_fast": {"resampling_method": TA_RESAMPLE_KAISER, "lowpass_filter_width": 16, "rolloff": 0.85,
                        "beta": 8.555504641634386},
        "kaiser_best": {"resampling_method": TA_RESAMPLE_KAISER, "lowpass_filter_width": 16, "rolloff": 0.9475937167399596,
                        "beta": 14.769656459379492},
    }
    assert method in params.keys(), f"method must be one of {list(params.keys())}"
    return params[method]


def resample_kernel(orig_freq: int, new_freq: int, lowpass_filter_width: int = 6, rolloff: float = 0.99,
                    resampling_method: str = TA_RESAMPLE_SINC, beta: Optional[float] = None):
    """Taps of torchaudio.functional.resample (torchaudio/functional/functional.py `_get_sinc_resample_kernel`, the
    reference's dependency for df.io.resample): (kernel float32 [new][2 * width + orig], width, orig, new) for the
    gcd-reduced rates, with torchaudio's own mix of float32 / float64 intermediate precision."""
    gcd = math.gcd(int(orig_freq), int(new_freq))
    og, nw = int(orig_freq) // gcd, int(new_freq) // gcd
    if lowpass_filter_width <= 0:
        raise ValueError("Low pass filter width should be positive.")
    base_freq = min(og, nw) * rolloff
    width = math.ceil(lowpass_filter_width * og / base_freq)
    idx = torch.arange(-width, width + og, dtype=torch.float64)[None, None] / og
    # (torchaudio divides the integer arange in the default float32 before adding the float64 index grid)
    t = torch.arange(0, -nw, -1)[:, None, None] / nw + idx
    t *= base_freq
    t = t.clamp_(-lowpass_filter_width, lowpass_filter_width)
    if resampling_method == TA_RESAMPLE_SINC:
        window = torch.cos(t * math.pi / lowpass_filter_width / 2) ** 2
    elif resampling_method == TA_RESAMPLE_KAISER:
        if beta is None:
            beta = 14.769656459379492
        beta_t = torch.tensor(float(beta))   # float32, like torchaudio
        window = torch.i0(beta_t * torch.sqrt(1 - (t / lowpass_filter_width) ** 2)) / torch.i0(beta_t)
    else:
        raise ValueError(f"Invalid resampling method: {resampling_method}")
    t *= math.pi
    scale = base_freq / og
    kernels = torch.where(t == 0, torch.tensor(1.0, dtype=torch.float64), t.sin() / t)
    kernels *= window * scale
    return kernels[:, 0].to(torch.float32).contiguous(), width, og, nw


def resample(audio: Tensor, orig_sr: int, new_sr: int, method="sinc_fast", device: int = 0) -> Tensor:
    """df/io.py:116-118: float32 [..., T] -> [..., ceil(new_sr * T / orig_sr)] (CPU tensor in, CPU tensor out; the
    convolution runs on the GPU)."""
    if int(orig_sr) == int(new_sr):
        return audio
    kern, width, og, nw = resample_kernel(orig_sr, new_sr, **get_resample_params(method))
    shape = audio.shape
    x = audio.detach().to("cpu", torch.float32).reshape(-1, shape[-1]).contiguous()
    c, t = x.shape
    t_out = int(math.ceil(nw * t / og))
    out = torch.empty((c, t_out), dtype=torch.float32)
    check(_lib.lib().dfb_resample_host(int(device), x.data_ptr(), c, t, kern.data_ptr(), og, nw, width, out.data_ptr(), t_out))
    return out.reshape(*shape[:-1], t_out)
