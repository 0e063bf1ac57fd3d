"""Drop-in for ``df.enhance`` (DeepFilterNet/df/enhance.py): ``init_df``, ``df_features``,
``enhance`` with the reference's signatures, argument meaning and return types.  The whole
``enhance()`` path (pad -> STFT -> features -> DNN -> mask + deep filter -> ISTFT -> crop) is ONE
C-ABI call (``dfb_enhance_host``) into the CUDA library; nothing is computed on the CPU.
"""
from __future__ import annotations

import logging
import os
from typing import Optional, Tuple, Union

import numpy as np
import torch
from torch import Tensor

from . import _lib
from ._lib import check
from .libdf import DF
from .model import DfNet, load_model

logger = logging.getLogger("deepfilternet_b200")

PRETRAINED_MODELS = ("DeepFilterNet", "DeepFilterNet2", "DeepFilterNet3")
DEFAULT_MODEL = "DeepFilterNet3"
_REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def model_search_dirs():
    d = []
    if os.environ.get("DFB_MODEL_DIR"):
        d.append(os.environ["DFB_MODEL_DIR"])
    d.append(os.path.join(_REPO, "models", "_ref"))
    return d


def get_model_basedir(m: Optional[str]) -> str:
    """enhance.py:92-98.  The reference downloads the default models; there is no network here, so
    pretrained names resolve to a local directory ($DFB_MODEL_DIR/<name> or models/_ref/<name>)."""
    if m is None:
        m = DEFAULT_MODEL
    if os.path.isdir(m):
        return m
    for base in model_search_dirs():
        cand = os.path.join(base, m)
        if os.path.isdir(cand):
            return cand
    return m


def init_df(
    model_base_dir: Optional[str] = None,
    post_filter: bool = False,
    log_level: str = "INFO",
    log_file: Optional[str] = "enhance.log",
    config_allow_defaults: bool = True,
    epoch: Union[str, int, None] = "best",
    default_model: str = DEFAULT_MODEL,
    mask_only: bool = False,
    device: int = 0,
) -> Tuple[DfNet, DF, str, int]:
    """enhance.py:101-187 -> (model, df_state, suffix, epoch)."""
    model_base_dir = get_model_basedir(model_base_dir or default_model)
    if not os.path.isdir(model_base_dir):
        raise NotADirectoryError("Base directory not found at {}".format(model_base_dir))
    logger.setLevel(getattr(logging, str(log_level).upper(), logging.INFO))
    if post_filter:
        raise NotImplementedError("post_filter=True: the post filter is outside the built hot path")
    if mask_only:
        raise NotImplementedError("mask_only=True is outside the built hot path")
    if epoch is None or (isinstance(epoch, str) and epoch.lower() == "none"):
        raise NotImplementedError("epoch='none' (random weights): use weights.random_state_dict + DfNet")
    model, df_state, ep = load_model(model_base_dir, epoch=epoch, device=device)
    suffix = os.path.basename(os.path.abspath(model_base_dir))
    logger.info("Running on device cuda:%d", device)
    logger.info("Model loaded")
    return model, df_state, suffix, ep


def df_features(audio: Tensor, df: DF, nb_df: int, device=None, alpha: Optional[float] = None
                ) -> Tuple[Tensor, Tensor, Tensor]:
    """enhance.py:190-203: audio f32 CPU [C,T] -> (spec [C,1,Tf,F,2], erb_feat [C,1,Tf,E],
    spec_feat [C,1,Tf,nb_df,2]); one fused device pass.  ``alpha`` defaults to the reference's
    ``get_norm_alpha()`` (df/utils.py:108-124) of the config the DF state was loaded with (``init_df`` /
    ``load_model`` stash it on the DF object), else to the value for norm_tau = 1."""
    if alpha is None:
        alpha = getattr(df, "norm_alpha", None)
    if alpha is None:
        from .config import ModelConfig
        alpha = ModelConfig(sr=df.sr(), hop_size=df.hop_size()).norm_alpha
    x = np.ascontiguousarray(audio.detach().cpu().numpy(), dtype=np.float32)
    if x.ndim != 2 or x.size == 0:
        raise RuntimeError("[df] Input array empty or not contiguous.")
    c, t = x.shape
    tf, f, e = t // df.hop_size(), df.fft_size() // 2 + 1, df.nb_erb()
    spec = np.empty((c, 1, tf, f, 2), dtype=np.float32)
    erb_feat = np.empty((c, 1, tf, e), dtype=np.float32)
    spec_feat = np.empty((c, 1, tf, nb_df, 2), dtype=np.float32)
    check(_lib.lib().dfb_features_host(df.handle, x.ctypes.data, c, t, int(nb_df), float(alpha),
                                       spec.ctypes.data, erb_feat.ctypes.data, spec_feat.ctypes.data))
    out = tuple(torch.from_numpy(a) for a in (spec, erb_feat, spec_feat))
    if device is not None:
        out = tuple(a.to(device) for a in out)
    return out


@torch.no_grad()
def enhance(model: DfNet, df_state: DF, audio: Tensor, pad: bool = True,
            atten_lim_db: Optional[float] = None, out: Optional[Tensor] = None) -> Tensor:
    """enhance.py:206-250: audio f32 CPU [C,T] @ model sr -> enhanced f32 CPU [C,T]
    (or [C, (T // hop) * hop], delayed by n_fft - hop, when ``pad`` is False).
    ``out`` (extension): optional preallocated (e.g. pinned) CPU tensor for the result."""
    model.eval()
    if audio.dim() != 2:
        raise ValueError("audio must have shape [C, T]")
    x = audio.detach().to("cpu", torch.float32).contiguous()
    c, t = x.shape
    out_len = int(_lib.lib().dfb_enhance_out_len(df_state.handle, t, 1 if pad else 0))
    if out is None:
        out = torch.empty((c, out_len), dtype=torch.float32)
    elif out.shape != (c, out_len) or out.dtype != torch.float32 or out.is_cuda or not out.is_contiguous():
        raise ValueError(f"out must be a contiguous float32 CPU tensor of shape {(c, out_len)}")
    lim = abs(float(atten_lim_db)) if atten_lim_db is not None else 0.0
    check(_lib.lib().dfb_enhance_host(model.handle, df_state.handle, x.data_ptr(), c, t, 1 if pad else 0,
                                      lim, out.data_ptr()))
    return out


@torch.no_grad()
def enhance_device(model: DfNet, df_state: DF, audio: Tensor, pad: bool = True,
                   atten_lim_db: Optional[float] = None, out: Optional[Tensor] = None) -> Tensor:
    """Device-resident variant of :func:`enhance`: ``audio`` is a CUDA tensor [B,T] on the model's
    device and the result stays there (asynchronous on the current stream)."""
    if not audio.is_cuda or audio.dtype != torch.float32 or not audio.is_contiguous() or audio.dim() != 2:
        raise ValueError("enhance_device expects a contiguous float32 CUDA tensor of shape [B, T]")
    if audio.device != model.cuda_device:
        raise ValueError(f"audio lives on {audio.device}, the model on {model.cuda_device}")
    b, t = audio.shape
    out_len = int(_lib.lib().dfb_enhance_out_len(df_state.handle, t, 1 if pad else 0))
    if out is None:
        out = torch.empty((b, out_len), dtype=torch.float32, device=audio.device)
    elif (out.shape != (b, out_len) or out.dtype != torch.float32 or not out.is_cuda or out.device != audio.device
          or not out.is_contiguous()):
        raise ValueError(f"out must be a contiguous float32 CUDA tensor of shape {(b, out_len)} on {audio.device}")
    lim = abs(float(atten_lim_db)) if atten_lim_db is not None else 0.0
    with torch.cuda.device(audio.device):
        stream = torch.cuda.current_stream(audio.device).cuda_stream
        check(_lib.lib().dfb_enhance(model.handle, df_state.handle, audio.data_ptr(), b, t, 1 if pad else 0,
                                     lim, out.data_ptr(), stream))
    return out
