"""Stage-by-stage parity report of the CUDA path against the CPU oracle (run on the GPU box):
    python tests/gpu_diag.py [--model DeepFilterNet3] [--random] > gpurun_out/diag.txt
Prints max-abs / RMS errors of every intermediate so that one GPU call localises a bug."""
import argparse
import ctypes as C
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import dfnet_oracle as O  # noqa: E402
import libdf_oracle as LO  # noqa: E402

from deepfilternet_b200 import DfNet, _lib, enhance, libdf  # noqa: E402
from deepfilternet_b200.config import ModelConfig, load_config  # noqa: E402
from deepfilternet_b200.model import find_checkpoint, load_state_dict_file  # noqa: E402
from deepfilternet_b200.weights import random_state_dict  # noqa: E402


def err(name, a, b):
    a = np.asarray(a, dtype=np.float64); b = np.asarray(b, dtype=np.float64)
    if a.shape != b.shape:
        print(f"{name:28s} SHAPE MISMATCH {a.shape} vs {b.shape}"); return
    d = a - b
    print(f"{name:28s} max|d| {np.abs(d).max():.3e}  rms {np.sqrt((d**2).mean()):.3e}  ref rms {np.sqrt((b**2).mean()):.3e}  nan {int(np.isnan(a).sum())}")


def fetch(model, name, shape):
    out = np.empty(shape, dtype=np.float32)
    n = _lib.lib().dfb_model_debug_fetch(model.handle, name.encode(), out.ctypes.data, out.size)
    assert n == out.size, (name, n, out.size)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", default="DeepFilterNet3")
    ap.add_argument("--random", action="store_true")
    ap.add_argument("--ll", action="store_true", help="random weights with the _ll topology")
    ap.add_argument("--B", type=int, default=3)
    ap.add_argument("--T", type=int, default=24000)
    a = ap.parse_args()
    mdir = os.path.join(ROOT, "models", "_ref", a.model)
    if a.ll:
        cfg = ModelConfig(model="deepfilternet3", conv_ch=64, conv_lookahead=0, df_lookahead=0, conv_kernel=(2, 3),
                          emb_hidden_dim=512, df_hidden_dim=512, emb_num_layers=3, df_num_layers=3, lin_groups=16,
                          enc_lin_groups=16, df_gru_skip="groupedlinear", df_pathway_kernel_size_t=5)
        sd = random_state_dict(cfg, seed=3)
    elif a.random or not os.path.isdir(mdir):
        cfg = load_config(os.path.join(mdir, "config.ini")) if os.path.isdir(mdir) else None
        if cfg is None:
            cfg = ModelConfig(model="deepfilternet3", conv_ch=64, conv_lookahead=2, df_lookahead=2, emb_num_layers=3,
                              df_num_layers=2, lin_groups=16, enc_lin_groups=32, df_gru_skip="groupedlinear",
                              df_pathway_kernel_size_t=5)
        sd = random_state_dict(cfg, seed=3)
    else:
        cfg = load_config(os.path.join(mdir, "config.ini"))
        sd = load_state_dict_file(find_checkpoint(os.path.join(mdir, "checkpoints"))[0])
    print("config:", cfg.model, "lookahead", cfg.conv_lookahead, cfg.df_lookahead, "kt", cfg.conv_kernel, "H", cfg.emb_hidden_dim)
    from tests_common import synth_audio
    audio = synth_audio(a.B, a.T, seed=5)
    st = libdf.DF(cfg.sr, cfg.fft_size, cfg.hop_size, cfg.nb_erb, cfg.min_nb_erb_freqs)
    ost = LO.DF(cfg.sr, cfg.fft_size, cfg.hop_size, cfg.nb_erb, cfg.min_nb_erb_freqs)
    x = np.ascontiguousarray(audio.numpy())
    # ---- DSP
    spec = st.analysis(x); ospec = ost.analysis(x)
    err("analysis.spec", spec.view(np.float32), ospec.view(np.float32))
    err("erb(db)", libdf.erb(ospec, st.erb_widths()), LO.erb(ospec, ost.erb_widths()))
    e = LO.erb(ospec, ost.erb_widths())
    err("erb_norm", libdf.erb_norm(e, 0.99), LO.erb_norm(e, 0.99))
    err("unit_norm", libdf.unit_norm(ospec[..., :96].copy(), 0.99).view(np.float32), LO.unit_norm(ospec[..., :96].copy(), 0.99).view(np.float32))
    err("synthesis", st.synthesis(ospec.copy()), ost.synthesis(ospec.copy()))
    err("erb_inv", libdf.erb_inv(e, st.erb_widths()), LO.erb_inv(e, ost.erb_widths()))
    # ---- full oracle
    out_o, aux = O.enhance(sd, cfg.as_dict(), audio, pad=True, return_all=True)
    from deepfilternet_b200.enhance import df_features
    xp = torch.nn.functional.pad(audio, (0, cfg.fft_size))
    sp, fe, fs = df_features(xp, st, cfg.nb_df, alpha=cfg.norm_alpha)
    err("features.spec", sp, aux["spec"]); err("features.erb", fe, aux["erb_feat"]); err("features.spec_feat", fs, aux["spec_feat"])
    model = DfNet(cfg, sd, st)
    # feed ORACLE features into the DNN to isolate it
    spec_e, m, lsnr, last = model(aux["spec"], aux["erb_feat"], aux["spec_feat"])
    # intermediates from the oracle
    cfgd = cfg.as_dict()
    fs_o = aux["spec_feat"].squeeze(1).permute(0, 3, 1, 2); fe_o = aux["erb_feat"]
    lc = cfg.conv_lookahead
    if lc > 0:
        fe_o = torch.nn.functional.pad(fe_o, (0, 0, -lc, lc)); fs_o = torch.nn.functional.pad(fs_o, (0, 0, -lc, lc))
    e0, e1, e2, e3, emb, c0, lsnr_o = O.encoder(sd, cfgd, fe_o, fs_o)
    B, T = e0.shape[0], e0.shape[2]
    cl = lambda t: t.permute(0, 2, 3, 1).contiguous().numpy()  # [B,C,T,F] -> [B,T,F,C]
    err("e0", fetch(model, "e0", cl(e0).shape), cl(e0))
    err("e1", fetch(model, "e1", cl(e1).shape), cl(e1))
    err("e2", fetch(model, "e2", cl(e2).shape), cl(e2))
    if not cfg.enc_concat:
        err("e3", fetch(model, "e3", cl(e3).shape), cl(e3))
    err("c0", fetch(model, "c0", cl(c0).shape), cl(c0))
    err("emb", fetch(model, "emb", tuple(emb.shape)), emb.numpy())
    err("lsnr", lsnr, lsnr_o)
    err("m", m, aux["m"])
    if cfg.model == "deepfilternet3":
        err("coefs", last.permute(0, 2, 3, 1, 4).reshape(aux["coefs"].shape), aux["coefs"])
    err("spec_e", spec_e, aux["spec_e"])
    # ---- end to end
    out = enhance(model, st, audio)
    err("enhance(pad=True)", out, out_o)
    err("enhance(pad=False)", enhance(model, st, audio, pad=False), O.enhance(sd, cfgd, audio, pad=False))
    err("enhance(atten 12dB)", enhance(model, st, audio, atten_lim_db=12.0), O.enhance(sd, cfgd, audio, atten_lim_db=12.0))
    print("launches", int(_lib.lib().dfb_kernel_launches()))


if __name__ == "__main__":
    main()
