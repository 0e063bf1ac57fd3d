"""Host emulation (torch, fp64) of the DEVICE graph of DeepFilterNet v1 (csrc/dfb_model.cu forward_v1): the same sequence of
kernels on the same packed tensors and channel-last layouts, each kernel restated in a few lines.  Checks on the CPU what the
GPU parity tests check again on the device: the weight packing, the gather tables and the folded shuffles of
deepfilternet_b200/weights.py:pack_state_dict_v1.  Test infrastructure only."""
from __future__ import annotations

import numpy as np
import torch


def _t(a):
    return torch.from_numpy(np.asarray(a)).to(torch.float64)


def _idx(a):
    return torch.from_numpy(np.ascontiguousarray(a).view(np.int32).astype(np.int64))


def conv_in(x, w, b, kt, la):
    """k_conv_in: x [B,T,F,Cin] -> [B,T,F,C]; out[t] = relu(b + sum_{j,df} w[j][df][cin][:] x[t - (kt-1) + j + la][f + df - 1][cin])"""
    B, T, F, Cin = x.shape
    w = _t(w).reshape(kt, 3, Cin, -1)
    out = torch.zeros(B, T, F, w.shape[-1], dtype=torch.float64) + _t(b)
    xp = torch.zeros(B, T + 2 * kt + 2 * la + 2, F + 2, Cin, dtype=torch.float64)
    off = kt + la
    xp[:, off:off + T, 1:F + 1] = x
    for j in range(kt):
        for df in range(3):
            t0 = off - (kt - 1) + j + la
            out += torch.einsum("btfc,cn->btfn", xp[:, t0:t0 + T, df:df + F], w[j, df])
    return torch.relu(out)


def dwpw(x, P, name, mode, kt, la=0, path=None):
    """k_dwpw: depthwise (kt x 3, causal shifted by la) -> 1x1 -> bias -> relu; mode S1 / S2 / T2; x [B,T,Fin,C]"""
    dw, pw, b = _t(P[name + ".dw"]).reshape(kt, 3, -1), _t(P[name + ".pw"]), _t(P[name + ".b"])
    if path is not None:
        x = x + path
    B, T, Fin, C = x.shape
    Fout = Fin if mode == "S1" else (Fin // 2 if mode == "S2" else Fin * 2)
    xp = torch.zeros(B, T + kt + la + 2, Fin + 3, C, dtype=torch.float64)   # time offset kt, freq offset 1
    xp[:, kt:kt + T, 1:Fin + 1] = x
    acc = torch.zeros(B, T, Fout, C, dtype=torch.float64)
    for j in range(kt):
        t0 = kt - (kt - 1) + j + la
        xt = xp[:, t0:t0 + T]
        for fo in range(Fout):
            if mode == "S1":
                taps = [(fo + df - 1, df) for df in range(3)]
            elif mode == "S2":
                taps = [(2 * fo + df - 1, df) for df in range(3)]
            elif fo % 2 == 0:
                taps = [(fo // 2, 1)]
            else:
                taps = [(fo // 2, 2), (fo // 2 + 1, 0)]
            for fi, df in taps:
                if 0 <= fi < Fin:
                    acc[:, :, fo] += xt[:, :, fi + 1] * dw[j, df]
    return torch.relu(acc @ pw + b)


def grouped_linear(x, w, bias, G):
    """k_grouped_linear: x [..., G * Ig], w [G][Ig][Hg]"""
    w = _t(w)
    w = w.reshape(G, -1, w.numel() // G // (x.shape[-1] // G))
    xs = x.reshape(*x.shape[:-1], G, -1)
    y = torch.einsum("...gi,gih->...gh", xs, w).reshape(*x.shape[:-1], -1)
    return y + _t(bias)


def gru_dense(x, P, base):
    """one dense GRU layer from the packed block-diagonal weights; x [B,T,H]"""
    w_ih, w_hh = _t(P[base + ".w_ih_t"]).T, _t(P[base + ".w_hh"])
    H = w_hh.shape[1]
    w_ih, w_hh = w_ih.reshape(3 * H, -1), w_hh.reshape(3 * H, H)
    b_ih, b_hh = _t(P[base + ".b_ih"]), _t(P[base + ".b_hh"])
    B, T, _ = x.shape
    h = torch.zeros(B, H, dtype=torch.float64)
    ys = []
    xp = x @ w_ih.T + b_ih
    for t in range(T):
        hp = h @ w_hh.T + b_hh
        r = torch.sigmoid(xp[:, t, :H] + hp[:, :H])
        z = torch.sigmoid(xp[:, t, H:2 * H] + hp[:, H:2 * H])
        n = torch.tanh(xp[:, t, 2 * H:] + r * hp[:, 2 * H:])
        h = (1 - z) * n + z * h
        ys.append(h)
    return torch.stack(ys, 1)


def forward(P: dict, d: dict, feat_erb: torch.Tensor, feat_spec: torch.Tensor):
    """feat_erb [B,T,E], feat_spec [B,T,Fd,2] -> m [B,T,E], coefs [B,T,Fd*2O] (device layout f * 2O + k), lsnr [B,T], alpha [B,T]"""
    B, T, E = feat_erb.shape
    Fd, C, kt, H = d["nb_df"], d["conv_ch"], d["conv_kt"], d["emb_hidden"]
    fe, fs = feat_erb.to(torch.float64).unsqueeze(-1), feat_spec.to(torch.float64)
    e0 = conv_in(fe, P["enc.erb_conv0.w"], P["enc.erb_conv0.b"], d["inp_kt"], 1)
    e1 = dwpw(e0, P, "enc.erb_conv1", "S2", kt, la=1)
    e2 = dwpw(e1, P, "enc.erb_conv2", "S2", kt)
    e3 = dwpw(e2, P, "enc.erb_conv3", "S1", kt)
    c0 = conv_in(fs, P["enc.df_conv0.w"], P["enc.df_conv0.b"], d["inp_kt"], d["conv_lookahead"])
    c1 = dwpw(c0, P, "enc.df_conv1", "S2", kt)
    c1g = c1.reshape(B, T, -1)[..., _idx(P["v1.idx_c1"])]
    cemb = grouped_linear(c1g, P["enc.df_fc_emb.gl"], P["enc.df_fc_emb.bias"], d["g_df_fc_emb"])
    emb = e3.reshape(B, T, -1)[..., _idx(P["v1.idx_e3"])] + cemb[..., _idx(P["v1.idx_shuf"])]
    gs = _idx(P["v1.idx_gshuf"])

    def ggru(x, name, layers):
        ys = []
        for l in range(layers):
            x = gru_dense(x, P, f"{name}.g{l}.l0")
            ys.append(x)
        return sum(y[..., gs] for y in ys[:-1]) + ys[-1]

    emb = ggru(emb, "enc.emb_gru", d["enc_gru_layers"])
    lsnr = torch.sigmoid(emb @ _t(P["enc.lsnr.w"]) + _t(P["enc.lsnr.b"])) * d["lsnr_scale"] + d["lsnr_offset"]
    dec = torch.relu(grouped_linear(emb, P["erb_dec.fc_emb.gl"], P["erb_dec.fc_emb.bias"], d["g_erb_in"]))
    dec = dec[..., _idx(P["v1.idx_dec"])].reshape(B, T, E // 4, C)
    p3, p2 = dwpw(e3, P, "erb_dec.conv3p", "S1", 1), dwpw(e2, P, "erb_dec.conv2p", "S1", 1)
    p1, p0 = dwpw(e1, P, "erb_dec.conv1p", "S1", 1), dwpw(e0, P, "erb_dec.conv0p", "S1", 1)
    d3 = dwpw(dec, P, "erb_dec.convt3", "S1", kt, path=p3)
    d2 = dwpw(d3, P, "erb_dec.convt2", "T2", kt, path=p2)
    d1 = dwpw(d2, P, "erb_dec.convt1", "T2", kt, path=p1)
    x = d1 + p0
    w = _t(P["erb_dec.conv0_out.w"]).reshape(kt, 3, C)
    xp = torch.zeros(B, T + kt, E + 2, C, dtype=torch.float64)
    xp[:, kt - 1:kt - 1 + T, 1:E + 1] = x
    m = torch.zeros(B, T, E, dtype=torch.float64) + _t(P["erb_dec.conv0_out.b"])
    for j in range(kt):
        for df in range(3):
            m += (xp[:, j:j + T, df:df + E] * w[j, df]).sum(-1)
    m = torch.sigmoid(m)
    c = ggru(emb, "df_dec.df_gru", d["df_gru_layers"])
    alpha = torch.sigmoid(c @ _t(P["df_dec.df_fc_a.w"]) + _t(P["df_dec.df_fc_a.b"]))
    O2 = 2 * d["df_order"]
    cp = torch.relu(c0 @ _t(P["df_dec.df_convp.w"]).reshape(C, O2) + _t(P["df_dec.df_convp.b"])).reshape(B, T, Fd * O2)
    coefs = torch.tanh(c @ _t(P["df_dec.df_fc_out.w_t"]).reshape(H, Fd * O2) + _t(P["df_dec.df_fc_out.b"])) + cp
    return m, coefs, lsnr, alpha
