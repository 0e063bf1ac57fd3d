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
from .io import load_audio, resample, save_audio
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
    if epoch is None or (isinstance(epoch, str) and epoch.lower() == "none"):
        raise NotImplementedError("epoch='none' (random weights): use weights.random_state_dict + DfNet")
    model, df_state, ep = load_model(model_base_dir, epoch=epoch, device=device, post_filter=post_filter, mask_only=mask_only)
    suffix = os.path.basename(os.path.abspath(model_base_dir))
    if post_filter:
        suffix += "_pf"   # enhance.py:183-184
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


# ------------------------------------------------------------------------------------------ CLI ----
def parse_epoch_type(value: str) -> Union[int, str]:
    """enhance.py:253-261."""
    try:
        return int(value)
    except ValueError:
        assert value in ("best", "latest")
        return value


def setup_df_argument_parser(default_log_level: str = "INFO", parser=None):
    """enhance.py:299-339 (same options)."""
    import argparse
    if parser is None:
        parser = argparse.ArgumentParser()
    parser.add_argument("--model-base-dir", "-m", type=str, default=None,
                        help="Model directory containing checkpoints and config, or a pretrained model name.")
    parser.add_argument("--pf", help="Post-filter that slightly over-attenuates very noisy sections.", action="store_true")
    parser.add_argument("--output-dir", "-o", type=str, default=None, help="Directory in which the enhanced audio files will be stored.")
    parser.add_argument("--log-level", type=str, default=default_log_level, help="Logger verbosity. Can be one of (debug, info, error, none)")
    parser.add_argument("--debug", "-d", action="store_const", const="DEBUG", dest="log_level")
    parser.add_argument("--epoch", "-e", default="best", type=parse_epoch_type,
                        help="Epoch for checkpoint loading. Can be one of ['best', 'latest', <int>].")
    return parser


def main(args) -> int:
    """The `deepFilter` command (enhance.py:47-89): load each file at the model rate, enhance, resample back to the
    file's rate, save next to it (or into --output-dir) with the model suffix."""
    import glob
    import time
    model, df_state, suffix, _ = init_df(args.model_base_dir, post_filter=args.pf, log_level=args.log_level,
                                         config_allow_defaults=True, epoch=args.epoch, mask_only=args.no_df_stage)
    suffix = suffix if args.suffix else None
    if args.output_dir is None:
        args.output_dir = "."
    elif not os.path.isdir(args.output_dir):
        os.mkdir(args.output_dir)
    df_sr = model.cfg.sr
    if args.noisy_dir is not None:
        if len(args.noisy_audio_files) > 0:
            logger.error("Only one of `noisy_audio_files` or `noisy_dir` arguments are supported.")
            return 1
        input_files = sorted(glob.glob(args.noisy_dir + "/*"))
    else:
        assert len(args.noisy_audio_files) > 0, "No audio files provided"
        input_files = args.noisy_audio_files
    n_samples = len(input_files)
    for i, file in enumerate(input_files):
        if not os.path.isfile(file):
            logger.warning("File not found: %s. Skipping...", file)
            continue
        audio, meta = load_audio(file, df_sr, verbose=False)
        progress = (i + 1) / n_samples * 100
        t0 = time.time()
        audio = enhance(model, df_state, audio, pad=args.compensate_delay, atten_lim_db=args.atten_lim)
        t = time.time() - t0
        t_audio = audio.shape[-1] / df_sr
        p_str = f"{progress:2.0f}% | " if n_samples > 1 else ""
        logger.info("%sEnhanced noisy audio file '%s' in %.2fs (RT factor: %.3f)", p_str, os.path.basename(file), t, t / t_audio)
        audio = resample(audio.to("cpu"), df_sr, meta.sample_rate)
        save_audio(file, audio, sr=meta.sample_rate, output_dir=args.output_dir, suffix=suffix, log=False)
    return 0


def run(argv=None) -> int:
    """enhance.py:342-379."""
    parser = setup_df_argument_parser()
    parser.add_argument("--no-delay-compensation", dest="compensate_delay", action="store_false",
                        help="Don't add some padding to compensate the delay introduced by the real-time STFT/ISTFT implementation.")
    parser.add_argument("--atten-lim", "-a", type=int, default=None,
                        help="Attenuation limit in dB by mixing the enhanced signal with the noisy signal.")
    parser.add_argument("noisy_audio_files", type=str, nargs="*", help="List of noisy files to enhance.")
    parser.add_argument("--noisy-dir", "-i", type=str, default=None,
                        help="Input directory containing noisy audio files. Use instead of `noisy_audio_files`.")
    parser.add_argument("--no-suffix", action="store_false", dest="suffix", help="Don't add the model suffix to the enhanced audio files")
    parser.add_argument("--no-df-stage", action="store_true")
    return main(parser.parse_args(argv))


if __name__ == "__main__":
    raise SystemExit(run())
