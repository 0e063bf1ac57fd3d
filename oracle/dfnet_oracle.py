"""CPU fp32 restatement (plain torch functional ops) of the reference's DeepFilterNet2 /
DeepFilterNet3 / DeepFilterNet3_ll forward pass.

TEST INFRASTRUCTURE ONLY (see oracle/libdf_oracle.c header): imported by tests/,
__graft_entry__.smoke() and bench.py's CPU-baseline leg; never by the product package.

It consumes a reference ``state_dict`` (tensor names as in the shipped checkpoints, SURVEY.md
Appendix B) and a plain ``dict`` of hyper-parameters and follows, line by line:

  * Conv2dNormAct / ConvTranspose2dNormAct   DeepFilterNet/df/modules.py:18-72, 75-126
  * GroupedLinearEinsum                       modules.py:741-780
  * SqueezedGRU / SqueezedGRU_S               modules.py:663-699, 702-738
  * Mask                                      modules.py:248-269
  * MF.DF                                     multiframe.py:72-74, 126-136, 169-180
  * Encoder / ErbDecoder / DfDecoder / DfNet  deepfilternet3.py:100-456, deepfilternet2.py:98-505

Pinned against the reference itself: tests/golden/dfnet_*.npz hold outputs of the reference
modules (imported in the build container by oracle/gen_golden.py) and tests/test_oracle_golden.py
compares this file with them.
"""
from __future__ import annotations

import math
from typing import Dict, Optional, Tuple

import torch
import torch.nn.functional as F
from torch import Tensor

SD = Dict[str, Tensor]

DEFAULTS_DFN3 = dict(
    model="deepfilternet3", sr=48000, fft_size=960, hop_size=480, nb_erb=32, nb_df=96,
    df_order=5, df_lookahead=2, conv_lookahead=2, conv_ch=64, conv_kernel=(1, 3),
    convt_kernel=(1, 3), conv_kernel_inp=(3, 3), emb_hidden_dim=256, emb_num_layers=3,
    df_hidden_dim=256, df_num_layers=2, lin_groups=16, enc_lin_groups=32, enc_concat=False,
    df_gru_skip="groupedlinear", df_pathway_kernel_size_t=5, lsnr_max=35, lsnr_min=-15,
)


def _seq_entries(sd: SD, prefix: str):
    """Sorted (index, kind) of the nn.Sequential children that own tensors."""
    idx = sorted({int(k[len(prefix) + 1:].split(".")[0]) for k in sd if k.startswith(prefix + ".")})
    out = []
    for i in idx:
        if f"{prefix}.{i}.running_mean" in sd:
            out.append((i, "bn"))
        else:
            out.append((i, "conv"))
    return out


def _bn(x: Tensor, sd: SD, p: str) -> Tensor:
    return F.batch_norm(x, sd[p + ".running_mean"], sd[p + ".running_var"], sd[p + ".weight"],
                        sd[p + ".bias"], training=False, eps=1e-5)


def conv_norm_act(x: Tensor, sd: SD, prefix: str, fstride: int = 1, act: str = "relu",
                  transposed: bool = False) -> Tensor:
    """modules.py:18-72 (Conv2dNormAct) / :75-126 (ConvTranspose2dNormAct); x: [B,C,T,F]."""
    entries = _seq_entries(sd, prefix)
    convs = [i for i, k in entries if k == "conv"]
    bns = [i for i, k in entries if k == "bn"]
    w = sd[f"{prefix}.{convs[0]}.weight"]
    kt, kf = w.shape[2], w.shape[3]
    if kt > 1:  # causal time padding, modules.py:45-48 / :100-103
        x = F.pad(x, (0, 0, kt - 1, 0))
    if not transposed:
        groups = x.shape[1] // w.shape[1]
        x = F.conv2d(x, w, None, stride=(1, fstride), padding=(0, kf // 2), groups=groups)
    else:
        out_ch = sd[f"{prefix}.{bns[0]}.weight"].shape[0]
        groups = out_ch // w.shape[1]  # ConvTranspose2d weight is [in, out/groups, kt, kf]
        x = F.conv_transpose2d(x, w, None, stride=(1, fstride), padding=(kt - 1, kf // 2),
                               output_padding=(0, kf // 2), groups=groups)
    if len(convs) > 1:  # separable: 1x1 pointwise, modules.py:66-67
        x = F.conv2d(x, sd[f"{prefix}.{convs[1]}.weight"])
    if bns:
        x = _bn(x, sd, f"{prefix}.{bns[0]}")
    if act == "relu":
        x = F.relu(x)
    elif act == "sigmoid":
        x = torch.sigmoid(x)
    return x


def grouped_linear(x: Tensor, w: Tensor) -> Tensor:
    """modules.py:766-776: x [B,T,I], w [G, I/G, H/G] -> [B,T,H]"""
    b, t, _ = x.shape
    g = w.shape[0]
    return torch.einsum("btgi,gih->btgh", x.view(b, t, g, -1), w).flatten(2, 3)


def gru(x: Tensor, sd: SD, prefix: str, num_layers: int) -> Tensor:
    """torch.nn.GRU(batch_first=True), h0 = 0 (modules.py:684,723)."""
    hidden = sd[f"{prefix}.weight_hh_l0"].shape[1]
    m = torch.nn.GRU(sd[f"{prefix}.weight_ih_l0"].shape[1], hidden, num_layers=num_layers,
                     batch_first=True)
    with torch.no_grad():
        for l in range(num_layers):
            for n in ("weight_ih", "weight_hh", "bias_ih", "bias_hh"):
                getattr(m, f"{n}_l{l}").copy_(sd[f"{prefix}.{n}_l{l}"])
    m.eval()
    with torch.no_grad():
        y, _ = m(x)
    return y


def _n_gru_layers(sd: SD, prefix: str) -> int:
    n = 0
    while f"{prefix}.weight_ih_l{n}" in sd:
        n += 1
    return n


def squeezed_gru_s(x: Tensor, sd: SD, prefix: str, skip: Optional[Tensor] = None) -> Tensor:
    """SqueezedGRU_S (DFN3), modules.py:732-738: lin_in+ReLU -> GRU -> [lin_out+ReLU] -> +skip(input)"""
    y = F.relu(grouped_linear(x, sd[f"{prefix}.linear_in.0.weight"]))
    y = gru(y, sd, f"{prefix}.gru", _n_gru_layers(sd, f"{prefix}.gru"))
    if f"{prefix}.linear_out.0.weight" in sd:
        y = F.relu(grouped_linear(y, sd[f"{prefix}.linear_out.0.weight"]))
    return y


def squeezed_gru(x: Tensor, sd: SD, prefix: str, identity_skip: bool) -> Tensor:
    """SqueezedGRU (DFN2), modules.py:693-699: x=lin_in(in); y=GRU(x); y+=skip(x); lin_out(y)"""
    xi = F.relu(grouped_linear(x, sd[f"{prefix}.linear_in.0.weight"]))
    y = gru(xi, sd, f"{prefix}.gru", _n_gru_layers(sd, f"{prefix}.gru"))
    if identity_skip:
        y = y + xi
    if f"{prefix}.linear_out.0.weight" in sd:
        y = F.relu(grouped_linear(y, sd[f"{prefix}.linear_out.0.weight"]))
    return y


def encoder(sd: SD, cfg: dict, feat_erb: Tensor, feat_spec: Tensor):
    """deepfilternet3.py:166-185 / deepfilternet2.py:165-184"""
    e0 = conv_norm_act(feat_erb, sd, "enc.erb_conv0")
    e1 = conv_norm_act(e0, sd, "enc.erb_conv1", fstride=2)
    e2 = conv_norm_act(e1, sd, "enc.erb_conv2", fstride=2)
    e3 = conv_norm_act(e2, sd, "enc.erb_conv3")
    c0 = conv_norm_act(feat_spec, sd, "enc.df_conv0")
    c1 = conv_norm_act(c0, sd, "enc.df_conv1", fstride=2)
    cemb = c1.permute(0, 2, 3, 1).flatten(2)
    cemb = F.relu(grouped_linear(cemb, sd["enc.df_fc_emb.0.weight"]))
    emb = e3.permute(0, 2, 3, 1).flatten(2)
    emb = torch.cat((emb, cemb), dim=-1) if cfg["enc_concat"] else emb + cemb
    if cfg["model"] == "deepfilternet2":
        emb = squeezed_gru(emb, sd, "enc.emb_gru", identity_skip=False)
    else:
        emb = squeezed_gru_s(emb, sd, "enc.emb_gru")
    lsnr = torch.sigmoid(F.linear(emb, sd["enc.lsnr_fc.0.weight"], sd["enc.lsnr_fc.0.bias"]))
    lsnr = lsnr * (cfg["lsnr_max"] - cfg["lsnr_min"]) + cfg["lsnr_min"]
    return e0, e1, e2, e3, emb, c0, lsnr


def erb_decoder(sd: SD, cfg: dict, emb, e3, e2, e1, e0) -> Tensor:
    """deepfilternet3.py:245-254 / deepfilternet2.py:248-258"""
    b, _, t, f8 = e3.shape
    if cfg["model"] == "deepfilternet2":
        emb = squeezed_gru(emb, sd, "erb_dec.emb_gru", identity_skip=True)
    else:
        emb = squeezed_gru_s(emb, sd, "erb_dec.emb_gru")
    emb = emb.view(b, t, f8, -1).permute(0, 3, 1, 2)
    e3 = conv_norm_act(conv_norm_act(e3, sd, "erb_dec.conv3p") + emb, sd, "erb_dec.convt3")
    e2 = conv_norm_act(conv_norm_act(e2, sd, "erb_dec.conv2p") + e3, sd, "erb_dec.convt2",
                       fstride=2, transposed=True)
    e1 = conv_norm_act(conv_norm_act(e1, sd, "erb_dec.conv1p") + e2, sd, "erb_dec.convt1",
                       fstride=2, transposed=True)
    m = conv_norm_act(conv_norm_act(e0, sd, "erb_dec.conv0p") + e1, sd, "erb_dec.conv0_out",
                      act="sigmoid")
    return m


def df_decoder(sd: SD, cfg: dict, emb: Tensor, c0: Tensor) -> Tensor:
    """deepfilternet3.py:323-331 / deepfilternet2.py:363-371 -> coefs [B,T,F,O*2]"""
    b, t, _ = emb.shape
    if cfg["model"] == "deepfilternet2":
        c = squeezed_gru(emb, sd, "df_dec.df_gru", identity_skip=True)
    else:
        c = squeezed_gru_s(emb, sd, "df_dec.df_gru")
    if "df_dec.df_skip.weight" in sd:
        c = c + grouped_linear(emb, sd["df_dec.df_skip.weight"])
    c0 = conv_norm_act(c0, sd, "df_dec.df_convp").permute(0, 2, 3, 1)
    c = torch.tanh(grouped_linear(c, sd["df_dec.df_out.0.weight"]))
    return c.view(b, t, cfg["nb_df"], cfg["df_order"] * 2) + c0


def erb_inv_matrix(widths, dtype=torch.float32) -> Tensor:
    """modules.py:206-223 with inverse=True, normalized=True: 0/1 matrix [E, F]."""
    f = int(sum(int(w) for w in widths))
    fb = torch.zeros((len(widths), f), dtype=dtype)
    o = 0
    for i, w in enumerate(widths):
        fb[i, o:o + int(w)] = 1
        o += int(w)
    return fb


def apply_mask(spec: Tensor, m: Tensor, erb_inv_fb: Tensor) -> Tensor:
    """modules.py:266-269"""
    return spec * m.matmul(erb_inv_fb).unsqueeze(4)


def deep_filter(spec: Tensor, coefs: Tensor, nb_df: int, order: int, lookahead: int) -> Tensor:
    """multiframe.py:169-180; spec [B,1,T,F,2] (returned modified copy), coefs [B,T,Fd,O*2]."""
    b, _, t, _, _ = spec.shape
    sc = torch.view_as_complex(spec.contiguous())  # [B,1,T,F]
    padded = F.pad(sc, (0, 0, order - 1 - lookahead, lookahead))
    unf = padded.unfold(2, order, 1)[..., :nb_df, :]  # [B,1,T,Fd,O]
    cc = torch.view_as_complex(coefs.reshape(b, t, nb_df, order, 2).contiguous())  # [B,T,Fd,O]
    y = torch.einsum("bctfn,btfn->bctf", unf, cc)
    out = spec.clone()
    out[..., :nb_df, :] = torch.view_as_real(y)
    return out


@torch.no_grad()
def dfnet_forward(sd: SD, cfg: dict, erb_widths, spec: Tensor, feat_erb: Tensor,
                  feat_spec: Tensor) -> Tuple[Tensor, Tensor, Tensor, Tensor]:
    """DfNet.forward: deepfilternet3.py:389-456 / deepfilternet2.py:481-505.

    spec [B,1,T,F,2], feat_erb [B,1,T,E], feat_spec [B,1,T,Fd,2]
    -> (spec_e [B,1,T,F,2], m [B,1,T,E], lsnr [B,T,1], coefs [B,T,Fd,O*2])
    """
    fs = feat_spec.squeeze(1).permute(0, 3, 1, 2)
    lc = cfg["conv_lookahead"]
    fe = feat_erb
    if lc > 0:  # ConstantPad2d((0,0,-lc,lc)), deepfilternet3.py:359,409-410
        fe = F.pad(fe, (0, 0, -lc, lc))
        fs = F.pad(fs, (0, 0, -lc, lc))
    e0, e1, e2, e3, emb, c0, lsnr = encoder(sd, cfg, fe, fs)
    m = erb_decoder(sd, cfg, emb, e3, e2, e1, e0)
    inv = erb_inv_matrix(erb_widths)
    pf, mask_only = bool(cfg.get("mask_pf", False)), bool(cfg.get("mask_only", False))
    m_app = m
    if pf and cfg["model"] == "deepfilternet2":
        # Mask.pf (modules.py:234-245, beta = 0.02): the post filter acts on the ERB gains
        beta = 0.02
        m_sin = m * torch.sin(math.pi * m / 2)
        m_app = (1 + beta) * m / (1 + beta * m.div(m_sin.clamp_min(1e-12)).pow(2))
    spec_m = apply_mask(spec, m_app, inv)
    coefs = df_decoder(sd, cfg, emb, c0)
    nb_df, order, la = cfg["nb_df"], cfg["df_order"], cfg["df_lookahead"]
    if cfg["model"] == "deepfilternet2":
        # deepfilternet2.py:494-503 (run_df = False with mask_only: checkpoint.py:32)
        spec_e = spec_m if mask_only else deep_filter(spec_m, coefs, nb_df, order, la)
    else:
        if mask_only:  # deepfilternet3.py:444-446
            spec_e = spec_m
        else:
            spec_e = deep_filter(spec, coefs, nb_df, order, la)  # deepfilternet3.py:442-443
            spec_e[..., nb_df:, :] = spec_m[..., nb_df:, :]
        if pf:  # deepfilternet3.py:448-454
            beta, eps = float(cfg.get("pf_beta", 0.02)), 1e-12
            mask = (torch.view_as_complex(spec_e.contiguous()).abs() / torch.view_as_complex(spec.contiguous()).abs().add(eps)).clamp(eps, 1)
            mask_sin = mask * torch.sin(math.pi * mask / 2).clamp_min(eps)
            g = (1 + beta) / (1 + beta * mask.div(mask_sin).pow(2))
            spec_e = spec_e * g.unsqueeze(-1)
    return spec_e, m, lsnr, coefs


def norm_alpha(sr: int, hop: int, tau: float) -> float:
    """df/utils.py:108-124"""
    import math
    a_ = math.exp(-hop / sr / tau)
    precision, a = 3, 1.0
    while a >= 1.0:
        a = round(a_, precision)
        precision += 1
    return a


@torch.no_grad()
def enhance(sd: SD, cfg: dict, audio: Tensor, pad: bool = True,
            atten_lim_db: Optional[float] = None, libdf=None, return_all: bool = False):
    """df/enhance.py:190-250 (df_features + enhance) on the CPU oracle."""
    import numpy as np
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
    spec_feat = torch.view_as_real(
        torch.as_tensor(libdf.unit_norm(np.ascontiguousarray(spec[..., :cfg["nb_df"]]), a))
    ).unsqueeze(1)
    spec_t = torch.view_as_real(torch.as_tensor(spec)).unsqueeze(1)
    spec_e, m, lsnr, coefs = dfnet_forward(sd, cfg, widths, spec_t.clone(), erb_feat, spec_feat)
    enh = torch.view_as_complex(spec_e.squeeze(1).contiguous())
    if atten_lim_db is not None and abs(atten_lim_db) > 0:
        lim = 10 ** (-abs(atten_lim_db) / 20)
        enh = torch.as_tensor(spec) * lim + enh * (1 - lim)
    out = torch.as_tensor(st.synthesis(np.ascontiguousarray(enh.numpy())))
    if pad:
        d = n_fft - hop
        out = out[:, d:orig_len + d]
    if return_all:
        return out, dict(spec=spec_t, erb_feat=erb_feat, spec_feat=spec_feat, spec_e=spec_e, m=m,
                         lsnr=lsnr, coefs=coefs)
    return out
