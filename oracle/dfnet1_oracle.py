"""CPU fp32 restatement (plain torch functional ops) of the reference's DeepFilterNet (v1) forward pass.

TEST INFRASTRUCTURE ONLY (see oracle/libdf_oracle.c header): imported by tests/ and bench.py's CPU-baseline
leg; never by the product package.

It consumes the shipped checkpoint's ``state_dict`` with the legacy ``clc_*`` keys already renamed to ``df_*``
(DeepFilterNet/df/checkpoint.py:78) and follows, line by line:

  * convkxf                                   DeepFilterNet/df/modules.py:129-193
  * GroupedLinear                             modules.py:783-813
  * GroupedGRULayer / GroupedGRU              modules.py:503-660
  * Mask                                      modules.py:248-269
  * DfOp.forward_real_unfold / assign_df      modules.py:388-406, 470-478
  * Encoder / ErbDecoder / DfDecoder / DfNet  deepfilternet.py:64-279

Pinned against the reference itself: tests/golden/dfnet1.npz holds outputs of the reference modules (imported in the
build container by oracle/gen_golden_v1.py) and tests/test_oracle_golden.py compares this file with them.
"""
from __future__ import annotations

from typing import Dict, Tuple

import torch
import torch.nn.functional as F
from torch import Tensor

SD = Dict[str, Tensor]

DEFAULTS_DFN1 = dict(
    model="deepfilternet", sr=48000, fft_size=960, hop_size=480, nb_erb=32, nb_df=96,
    df_order=5, df_lookahead=1, conv_lookahead=2, conv_ch=64, conv_k_enc=2, conv_k_dec=2,
    emb_hidden_dim=512, emb_num_layers=3, df_hidden_dim=512, df_num_layers=2, gru_groups=8,
    lin_groups=8, group_shuffle=True, lsnr_max=35, lsnr_min=-15,
)


def rename_legacy_keys(sd: SD) -> SD:
    """checkpoint.py:78: the shipped v1 checkpoint still calls deep filtering 'clc'."""
    return {k.replace("clc", "df"): v for k, v in sd.items()}


def convkxf(x: Tensor, sd: SD, p: str, k: int, fstride: int = 2, lookahead: int = 0, act: str = "relu",
            mode: str = "normal") -> Tensor:
    """modules.py:129-193; x [B,C,T,F].  The layer's structure is read off the tensors that exist: `sconv` / `sconvt`,
    optional `1x1conv` (groups > 1), optional `norm` (otherwise the conv carries a bias)."""
    # modules.py:151-154: time padding (k - 1 - lookahead, lookahead); a negative amount crops
    if k - 1 - lookahead != 0 or lookahead != 0:
        x = F.pad(x, (0, 0, k - 1 - lookahead, lookahead))
    if mode == "normal":
        w = sd[p + ".sconv.weight"]
        f = w.shape[3]
        stride = 1 if f == 1 else (1, fstride)
        x = F.conv2d(x, w, sd.get(p + ".sconv.bias"), stride=stride, padding=(0, (f - 1) // 2), groups=x.shape[1] // w.shape[1])
    else:  # "transposed", modules.py:172-181
        w = sd[p + ".sconvt.weight"]       # [in, out / groups, k, f]
        f = w.shape[3]
        out_ch = sd[p + ".norm.weight"].shape[0]
        x = F.conv_transpose2d(x, w, sd.get(p + ".sconvt.bias"), stride=(1, fstride), padding=(k - 1, (f - 1) // 2),
                               output_padding=(0, (f - 1) // 2), groups=out_ch // w.shape[1])
    if p + ".1x1conv.weight" in sd:
        x = F.conv2d(x, sd[p + ".1x1conv.weight"])
    if p + ".norm.weight" in sd:
        x = F.batch_norm(x, sd[p + ".norm.running_mean"], sd[p + ".norm.running_var"], sd[p + ".norm.weight"],
                         sd[p + ".norm.bias"], training=False, eps=1e-5)
    return F.relu(x) if act == "relu" else torch.sigmoid(x)


def grouped_linear(x: Tensor, sd: SD, p: str, groups: int, shuffle: bool) -> Tensor:
    """modules.py:803-813: `groups` nn.Linear on contiguous input slices, outputs concatenated then interleaved."""
    isz = x.shape[-1] // groups
    outs = [F.linear(x[..., i * isz:(i + 1) * isz], sd[f"{p}.layers.{i}.weight"], sd[f"{p}.layers.{i}.bias"]) for i in range(groups)]
    y = torch.cat(outs, dim=-1)
    if shuffle and groups > 1:
        hs = y.shape[-1] // groups
        y = y.view(-1, hs, groups).transpose(-1, -2).reshape(y.shape)
    return y


def _gru(x: Tensor, sd: SD, p: str) -> Tensor:
    """one nn.GRU(I, H), time major [T,B,I], h0 = 0"""
    w_ih, w_hh = sd[p + ".weight_ih_l0"], sd[p + ".weight_hh_l0"]
    m = torch.nn.GRU(w_ih.shape[1], w_hh.shape[1])
    with torch.no_grad():
        m.weight_ih_l0.copy_(w_ih); m.weight_hh_l0.copy_(w_hh)
        m.bias_ih_l0.copy_(sd[p + ".bias_ih_l0"]); m.bias_hh_l0.copy_(sd[p + ".bias_hh_l0"])
    m.eval()
    with torch.no_grad():
        return m(x)[0]


def grouped_gru(x: Tensor, sd: SD, p: str, layers: int, groups: int, shuffle: bool, add_outputs: bool = True) -> Tensor:
    """modules.py:636-660 (GroupedGRU.forward) over :556-574 (GroupedGRULayer.forward); x [T,B,I]."""
    t, b, _ = x.shape
    out = None
    for i in range(layers):
        isz = x.shape[-1] // groups
        x = torch.cat([_gru(x[..., g * isz:(g + 1) * isz], sd, f"{p}.grus.{i}.layers.{g}") for g in range(groups)], dim=-1)
        if shuffle and groups > 1 and i < layers - 1:
            x = x.view(t, b, -1, groups).transpose(2, 3).reshape(t, b, -1)
        out = x if (out is None or not add_outputs) else out + x
    return out


def erb_inv_matrix(widths, dtype=torch.float32) -> Tensor:
    """modules.py:206-223 with inverse=True, normalized=True: 0/1 matrix [E, F]."""
    f = int(sum(int(w) for w in widths))
    fb = torch.zeros((len(widths), f), dtype=dtype)
    o = 0
    for i, w in enumerate(widths):
        fb[i, o:o + int(w)] = 1
        o += int(w)
    return fb


def df_op_real_unfold(spec: Tensor, coefs: Tensor, alpha: Tensor, nb_df: int, order: int, lookahead: int) -> Tensor:
    """modules.py:388-406 + assign_df :470-478; spec [B,1,T,F,2], coefs [B,T,O,Fd,2], alpha [B,T,1]."""
    b = spec.shape[0]
    x = spec[..., :nb_df, :].squeeze(1)                                # [B,T,Fd,2]
    padded = F.pad(x, (0, 0, 0, 0, order - lookahead - 1, lookahead))  # spec_pad, dim = -3
    padded = padded.unfold(1, order, 1).permute(0, 1, 4, 2, 3)         # [B,T,O,Fd,2]
    re = padded[..., 0] * coefs[..., 0] - padded[..., 1] * coefs[..., 1]
    im = padded[..., 1] * coefs[..., 0] + padded[..., 0] * coefs[..., 1]
    spec_f = torch.stack((re, im), -1).sum(dim=2).unsqueeze(1)         # [B,1,T,Fd,2]
    out = spec.clone()
    a = alpha.view(b, 1, -1, 1, 1)
    out[..., :nb_df, :] = spec_f * a + spec[..., :nb_df, :] * (1 - a)
    return out


@torch.no_grad()
def dfnet1_forward(sd: SD, cfg: dict, erb_widths, spec: Tensor, feat_erb: Tensor, feat_spec: Tensor, run_df: bool = True):
    """DfNet.forward, deepfilternet.py:262-279.

    spec [B,1,T,F,2], feat_erb [B,1,T,E], feat_spec [B,1,T,Fd,2]
    -> (spec_e [B,1,T,F,2], m [B,1,T,E], lsnr [B,T,1], coefs [B,T,O,Fd,2], alpha [B,T,1])
    """
    ke, kd, cl = cfg["conv_k_enc"], cfg["conv_k_dec"], cfg["conv_lookahead"]
    G, LG, shuffle = cfg["gru_groups"], cfg["lin_groups"], cfg["group_shuffle"]
    k0 = 1 if ke == 1 and cl == 0 else max(2, ke)
    fs = feat_spec.transpose(1, 4).squeeze(4)            # [B,2,T,Fd]
    # ---- Encoder.forward, deepfilternet.py:122-141
    e0 = convkxf(feat_erb, sd, "enc.erb_conv0", k0, fstride=1, lookahead=1 if cl > 0 else 0)
    e1 = convkxf(e0, sd, "enc.erb_conv1", ke, lookahead=1 if cl > 1 else 0)
    e2 = convkxf(e1, sd, "enc.erb_conv2", ke, lookahead=1 if cl > 2 else 0)
    e3 = convkxf(e2, sd, "enc.erb_conv3", ke, fstride=1)
    c0 = convkxf(fs, sd, "enc.df_conv0", k0, fstride=1, lookahead=cl)
    c1 = convkxf(c0, sd, "enc.df_conv1", ke)
    b, _, t, _ = feat_erb.shape
    cemb = c1.permute(2, 0, 1, 3).reshape(t, b, -1)
    cemb = grouped_linear(cemb, sd, "enc.df_fc_emb", LG, True)          # GroupedLinear default shuffle=True (:90-92)
    emb = e3.permute(2, 0, 1, 3).reshape(t, b, -1) + cemb
    emb = grouped_gru(emb, sd, "enc.emb_gru", cfg["emb_num_layers"], G, shuffle).transpose(0, 1)   # [B,T,H]
    lsnr = torch.sigmoid(F.linear(emb, sd["enc.lsnr_fc.0.weight"], sd["enc.lsnr_fc.0.bias"]))
    lsnr = lsnr * (cfg["lsnr_max"] - cfg["lsnr_min"]) + cfg["lsnr_min"]
    # ---- ErbDecoder.forward, deepfilternet.py:179-189
    f8 = e3.shape[3]
    d = F.relu(grouped_linear(emb, sd, "erb_dec.fc_emb.0", LG, shuffle))
    d = d.view(b, t, -1, f8).transpose(1, 2)
    d3 = convkxf(convkxf(e3, sd, "erb_dec.conv3p", 1) + d, sd, "erb_dec.convt3", kd, fstride=1)
    d2 = convkxf(convkxf(e2, sd, "erb_dec.conv2p", 1) + d3, sd, "erb_dec.convt2", kd, mode="transposed")
    d1 = convkxf(convkxf(e1, sd, "erb_dec.conv1p", 1) + d2, sd, "erb_dec.convt1", kd, mode="transposed")
    m = convkxf(convkxf(e0, sd, "erb_dec.conv0p", 1) + d1, sd, "erb_dec.conv0_out", kd, fstride=1, act="sigmoid")
    spec_m = spec * m.matmul(erb_inv_matrix(erb_widths)).unsqueeze(4)   # Mask, modules.py:266-269
    # ---- DfDecoder.forward, deepfilternet.py:219-229
    nb_df, order = cfg["nb_df"], cfg["df_order"]
    c = grouped_gru(emb.transpose(0, 1), sd, "df_dec.df_gru", cfg["df_num_layers"], G, shuffle).transpose(0, 1)
    cp = convkxf(c0, sd, "df_dec.df_convp", 1).transpose(1, 2)          # [B,T,O*2,Fd]
    alpha = torch.sigmoid(F.linear(c, sd["df_dec.df_fc_a.0.weight"], sd["df_dec.df_fc_a.0.bias"]))
    co = torch.tanh(F.linear(c, sd["df_dec.df_fc_out.0.weight"], sd["df_dec.df_fc_out.0.bias"]))
    co = co.view(b, t, order * 2, nb_df).add(cp).view(b, t, order, 2, nb_df).transpose(3, 4)   # [B,T,O,Fd,2]
    if run_df:
        spec_e = df_op_real_unfold(spec_m, co, alpha, nb_df, order, cfg["df_lookahead"])
    else:
        spec_e, alpha = spec_m, torch.zeros(b, t, 1)
    return spec_e, m, lsnr, co, alpha


@torch.no_grad()
def enhance(sd: SD, cfg: dict, audio: Tensor, pad: bool = True, atten_lim_db=None, libdf=None, return_all: bool = False):
    """df/enhance.py:190-250 (df_features + enhance) for the v1 model on the CPU oracle."""
    import numpy as np
    from dfnet_oracle import norm_alpha
    if libdf is None:
        import libdf_oracle as libdf
    n_fft, hop = cfg["fft_size"], cfg["hop_size"]
    st = libdf.DF(cfg["sr"], n_fft, hop, cfg["nb_erb"], cfg.get("min_nb_erb_freqs", 2))
    orig_len = audio.shape[-1]
    if pad:
        audio = F.pad(audio, (0, n_fft))
    a = norm_alpha(cfg["sr"], hop, cfg.get("norm_tau", 1.0))
    spec = st.analysis(np.ascontiguousarray(audio.numpy()))
    widths = st.erb_widths()
    erb_feat = torch.as_tensor(libdf.erb_norm(libdf.erb(spec, widths), a)).unsqueeze(1)
    spec_feat = torch.view_as_real(torch.as_tensor(libdf.unit_norm(np.ascontiguousarray(spec[..., :cfg["nb_df"]]), a))).unsqueeze(1)
    spec_t = torch.view_as_real(torch.as_tensor(spec)).unsqueeze(1)
    spec_e, m, lsnr, coefs, alpha = dfnet1_forward(sd, cfg, widths, spec_t.clone(), erb_feat, spec_feat)
    enh = torch.view_as_complex(spec_e.squeeze(1).contiguous())
    if atten_lim_db is not None and abs(atten_lim_db) > 0:
        lim = 10 ** (-abs(atten_lim_db) / 20)
        enh = torch.as_tensor(spec) * lim + enh * (1 - lim)
    out = torch.as_tensor(st.synthesis(np.ascontiguousarray(enh.numpy())))
    if pad:
        d = n_fft - hop
        out = out[:, d:orig_len + d]
    if return_all:
        return out, dict(spec=spec_t, erb_feat=erb_feat, spec_feat=spec_feat, spec_e=spec_e, m=m, lsnr=lsnr, coefs=coefs, alpha=alpha)
    return out
