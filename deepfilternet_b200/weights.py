"""Weight packer: reference ``state_dict`` -> the flat fp32 tensors libdfb200 consumes.

Source tensor names/shapes: SURVEY.md Appendix B (dumped from the shipped checkpoints of
``DeepFilterNet/df/deepfilternet3.py`` / ``deepfilternet2.py``).  All activations on the device are
channel-last ``[B, T, F, C=64]``, so conv weights are stored tap-major with the channel fastest.
Eval-mode BatchNorm (eps 1e-5, modules.py:68-69) is folded into the bias-free conv in front of it:
``scale = gamma / sqrt(var + eps)``, ``w' = w * scale[out]``, ``b' = beta - mean * scale``.

Packed tensors (name -> layout):
  enc.erb_conv0.w   [kt][3][C]        dense 1->C conv (kt = conv_kernel_inp[0]), BN folded
  enc.erb_conv0.b   [C]
  <blk>.dw          [kt][3][C]        depthwise taps of erb_conv1-3, df_conv1, convt3 (conv) and
                                      convt2, convt1 (ConvTranspose2d taps as stored, kt = 1)
  <blk>.pw          [C_in][C_out]     1x1 conv, transposed, BN folded
  <blk>.b           [C]
  enc.df_conv0.w    [kt][3][2][C]     grouped 2->C conv composed with its 1x1 conv and BN (direct K = 18 conv)
  enc.df_conv0.dw   [kt][3][C]        (unfused form, kept for reference) 2->C grouped conv taps
  enc.df_conv0.pw/.b
  erb_dec.conv{3,2,1,0}p.s / .b  [C]  depthwise 1x1 + BN folded: relu(x * s + b)
  erb_dec.conv0_out.w [kt][3][C]      C->1 conv, BN folded;  erb_dec.conv0_out.b [1]
  df_dec.df_convp.w1 [kt5][10][C/2]   grouped (2) temporal conv: out o reads channels (o // 5) * C/2 + c
  df_dec.df_convp.w2 [10][10]         1x1 (in, out), BN folded;  df_dec.df_convp.b [10]
  *.gl              [G][I/G][H/G]     GroupedLinearEinsum weight as stored (modules.py:752-757)
  *.gl_bx           BF16 hi | lo image of the same weight in the tcgen05 kernel's operand layout (gl_bx_image)
  <gru>.l{n}.w_ih_t [I][3H] (transposed for the projection GEMM), .w_hh [3H][H], .b_ih [3H],
                    .b_hh [3H]        torch.nn.GRU gate order (r,z,n)
  enc.lsnr.w [emb_out], enc.lsnr.b [1]
"""
from __future__ import annotations

from typing import Dict, List, Tuple

import numpy as np
import torch

from .config import ModelConfig

EPS = 1e-5


def _np(t) -> np.ndarray:
    return np.ascontiguousarray(t.detach().cpu().to(torch.float32).numpy())


def _seq(sd, prefix: str) -> Tuple[List[int], List[int]]:
    idx = sorted({int(k[len(prefix) + 1:].split(".")[0]) for k in sd if k.startswith(prefix + ".")})
    convs = [i for i in idx if f"{prefix}.{i}.running_mean" not in sd]
    bns = [i for i in idx if f"{prefix}.{i}.running_mean" in sd]
    return convs, bns


def _bn_fold(sd, p: str) -> Tuple[np.ndarray, np.ndarray]:
    g, b = _np(sd[p + ".weight"]).astype(np.float64), _np(sd[p + ".bias"]).astype(np.float64)
    mu, var = _np(sd[p + ".running_mean"]).astype(np.float64), _np(sd[p + ".running_var"]).astype(np.float64)
    scale = g / np.sqrt(var + EPS)
    return scale, b - mu * scale


def bf16_planes(w: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
    """fp32 -> (hi, lo) BF16 planes with w ~= hi + lo (both round-to-nearest-even), each returned as a
    float32 array holding two BF16 values per element (the C ABI moves fp32 tensors only)."""
    t = torch.from_numpy(np.ascontiguousarray(w, dtype=np.float32))
    hi = t.to(torch.bfloat16)
    lo = (t - hi.to(torch.float32)).to(torch.bfloat16)
    pack = lambda b: np.ascontiguousarray(b.contiguous().view(torch.int16).numpy().view(np.float32))
    return pack(hi), pack(lo)


def umma_sw128_image(w_nk: np.ndarray) -> np.ndarray:
    """[N][64] fp32 -> the shared-memory image of the K-major, 128-byte-swizzled tcgen05 B operand: BF16 hi plane
    then lo plane, each N rows of 128 bytes with 16-byte chunk j of row n stored at chunk j ^ (n & 7)."""
    hi, lo = bf16_planes(w_nk)
    n = w_nk.shape[0]
    assert w_nk.shape[1] == 64 and n % 8 == 0
    planes = []
    for pl in (hi, lo):
        c = pl.reshape(n, 8, 4)  # [row][16-byte chunk][4 floats = 8 bf16]
        o = np.empty_like(c)
        rows = np.arange(n)
        for j in range(8):
            o[rows, j ^ (rows & 7)] = c[rows, j]
        planes.append(o.reshape(-1))
    return np.ascontiguousarray(np.concatenate(planes))


def gl_bx_image(w: np.ndarray) -> np.ndarray:
    """GroupedLinearEinsum weight [G][Ig][Hg] -> the shared-memory image of the tcgen05 grouped-linear kernel's B operand
    (csrc/dfb_gl.cu): BF16 hi plane then lo plane, each [G][Ig/8][Hgp/8][8 n][8 i] -- K-major 8 x 16-byte core matrices,
    Hg zero-padded to a multiple of 16 (Hgp).  Returned as float32 words (two BF16 per element)."""
    G, Ig, Hg = w.shape
    assert Ig % 8 == 0
    Hgp = (Hg + 15) // 16 * 16
    wp = np.zeros((G, Ig, Hgp), dtype=np.float32)
    wp[:, :, :Hg] = w
    t = torch.from_numpy(wp)
    hi = t.to(torch.bfloat16)
    lo = (t - hi.to(torch.float32)).to(torch.bfloat16)
    planes = []
    for pl in (hi, lo):
        a = pl.view(torch.int16).numpy().reshape(G, Ig // 8, 8, Hgp // 8, 8)   # [g][kc][i8][rg][n8]
        planes.append(np.ascontiguousarray(a.transpose(0, 1, 3, 4, 2)).reshape(-1))  # [g][kc][rg][n8][i8]
    return np.ascontiguousarray(np.concatenate(planes)).view(np.float32)


def gru_layers(sd, prefix: str) -> int:
    n = 0
    while f"{prefix}.weight_ih_l{n}" in sd:
        n += 1
    return n


def pack_state_dict(sd: Dict[str, torch.Tensor], cfg: ModelConfig) -> Tuple[Dict[str, np.ndarray], dict]:
    """Returns (packed tensors, derived integer config for dfb_model_config)."""
    C = cfg.conv_ch
    out: Dict[str, np.ndarray] = {}

    def f32(a):
        return np.ascontiguousarray(np.asarray(a, dtype=np.float32))

    # --- erb_conv0: dense 1->C (kt,3) + BN
    convs, bns = _seq(sd, "enc.erb_conv0")
    w = _np(sd[f"enc.erb_conv0.{convs[0]}.weight"]).astype(np.float64)  # [C,1,kt,3]
    s, b = _bn_fold(sd, f"enc.erb_conv0.{bns[0]}")
    inp_kt = w.shape[2]
    out["enc.erb_conv0.w"] = f32((w[:, 0] * s[:, None, None]).transpose(1, 2, 0))
    out["enc.erb_conv0.b"] = f32(b)

    def dwpw(prefix: str, transposed: bool = False):
        convs, bns = _seq(sd, prefix)
        dw = _np(sd[f"{prefix}.{convs[0]}.weight"]).astype(np.float64)  # [C,1,kt,3] (conv and convT alike)
        assert dw.shape[0] == C and dw.shape[1] == 1 and dw.shape[3] == 3, (prefix, dw.shape)
        pw = _np(sd[f"{prefix}.{convs[1]}.weight"]).astype(np.float64)[:, :, 0, 0]  # [out,in]
        s, b = _bn_fold(sd, f"{prefix}.{bns[0]}")
        out[prefix + ".dw"] = f32(dw[:, 0].transpose(1, 2, 0))
        out[prefix + ".pw"] = f32((pw * s[:, None]).T)
        out[prefix + ".pw_nk"] = f32(pw * s[:, None])  # [C_out][C_in]: B operand of the tcgen05 kernel
        out[prefix + ".pw_sw"] = umma_sw128_image(out[prefix + ".pw_nk"])  # BF16x3 tcgen05 path
        out[prefix + ".b"] = f32(b)
        return dw.shape[2]

    kts = [dwpw(p) for p in ("enc.erb_conv1", "enc.erb_conv2", "enc.erb_conv3", "enc.df_conv1",
                             "erb_dec.convt3")]
    conv_kt = kts[0]
    assert all(k == conv_kt for k in kts)
    for p in ("erb_dec.convt2", "erb_dec.convt1"):
        assert dwpw(p, transposed=True) == 1, "ConvTranspose2d with kt > 1 not supported"
    # --- df_conv0: 2->C groups=2 (kt,3), 1x1, BN
    convs, bns = _seq(sd, "enc.df_conv0")
    dw = _np(sd[f"enc.df_conv0.{convs[0]}.weight"]).astype(np.float64)  # [C,1,kt,3]
    assert dw.shape[2] == inp_kt
    pw = _np(sd[f"enc.df_conv0.{convs[1]}.weight"]).astype(np.float64)[:, :, 0, 0]
    s, b = _bn_fold(sd, f"enc.df_conv0.{bns[0]}")
    # composed weights of the direct 2 -> C conv: W[dt][df][ri][n] = sum_{c in group ri} dw[c][dt][df] * pw'[n][c]
    g = C // 2
    pws = pw * s[:, None]                                 # [n][c], BN folded
    weff = np.stack([np.einsum("ctf,nc->tfn", dw[ri * g:(ri + 1) * g, 0], pws[:, ri * g:(ri + 1) * g]) for ri in range(2)], axis=2)
    out["enc.df_conv0.w"] = f32(weff)                     # [kt][3][2][C]
    out["enc.df_conv0.dw"] = f32(dw[:, 0].transpose(1, 2, 0))
    out["enc.df_conv0.pw"] = f32((pw * s[:, None]).T)
    out["enc.df_conv0.pw_nk"] = f32(pw * s[:, None])
    out["enc.df_conv0.b"] = f32(b)
    # --- decoder pathway convs (depthwise 1x1 + BN + ReLU)
    for n in (3, 2, 1, 0):
        p = f"erb_dec.conv{n}p"
        convs, bns = _seq(sd, p)
        w = _np(sd[f"{p}.{convs[0]}.weight"]).astype(np.float64)[:, 0, 0, 0]
        s, b = _bn_fold(sd, f"{p}.{bns[0]}")
        out[p + ".s"] = f32(w * s)
        out[p + ".b"] = f32(b)
    # --- conv0_out: C->1 (kt,3) + BN + sigmoid
    convs, bns = _seq(sd, "erb_dec.conv0_out")
    w = _np(sd[f"erb_dec.conv0_out.{convs[0]}.weight"]).astype(np.float64)  # [1,C,kt,3]
    s, b = _bn_fold(sd, f"erb_dec.conv0_out.{bns[0]}")
    assert w.shape[2] == conv_kt
    out["erb_dec.conv0_out.w"] = f32((w[0] * s[0]).transpose(1, 2, 0))
    out["erb_dec.conv0_out.b"] = f32(b)
    # --- df_convp: C->10 groups=2 (kt5,1), 1x1 10x10, BN, ReLU
    convs, bns = _seq(sd, "df_dec.df_convp")
    w1 = _np(sd[f"df_dec.df_convp.{convs[0]}.weight"]).astype(np.float64)  # [10, C/2, kt5, 1]
    w2 = _np(sd[f"df_dec.df_convp.{convs[1]}.weight"]).astype(np.float64)[:, :, 0, 0]  # [out,in]
    s, b = _bn_fold(sd, f"df_dec.df_convp.{bns[0]}")
    assert w1.shape[0] == 2 * cfg.df_order and w1.shape[1] == C // 2
    out["df_dec.df_convp.w1"] = f32(w1[:, :, :, 0].transpose(2, 0, 1))
    # tensor-core form (k_df_convp_tc): W2[n = g*32 + dt*O + o][k = channel] = w1[g*O + o][k - 32 g][dt] inside group g, else 0
    O, ktp = cfg.df_order, w1.shape[2]
    if C == 64 and ktp * O <= 32:
        w2x = np.zeros((64, 64), dtype=np.float32)
        for g_ in range(2):
            for dt in range(ktp):
                for o in range(O):
                    w2x[g_ * 32 + dt * O + o, g_ * 32:(g_ + 1) * 32] = w1[g_ * O + o, :, dt, 0]
        out["df_dec.df_convp.w_sw"] = umma_sw128_image(w2x)
    out["df_dec.df_convp.w2"] = f32((w2 * s[:, None]).T)
    out["df_dec.df_convp.b"] = f32(b)
    pathway_kt = w1.shape[2]

    # --- grouped linears
    def gl(dst: str, src: str) -> int:
        w = _np(sd[src])
        out[dst] = f32(w)
        if w.shape[1] % 16 == 0 and w.shape[2] % 4 == 0:
            out[dst + "_bx"] = gl_bx_image(w)   # tcgen05 grouped linear (BF16x3)
        return w.shape[0]

    g = {}
    g["g_df_fc_emb"] = gl("enc.df_fc_emb.gl", "enc.df_fc_emb.0.weight")
    g["g_enc_in"] = gl("enc.emb_gru.in.gl", "enc.emb_gru.linear_in.0.weight")
    g["g_enc_out"] = gl("enc.emb_gru.out.gl", "enc.emb_gru.linear_out.0.weight") \
        if "enc.emb_gru.linear_out.0.weight" in sd else 0
    g["g_erb_in"] = gl("erb_dec.emb_gru.in.gl", "erb_dec.emb_gru.linear_in.0.weight")
    g["g_erb_out"] = gl("erb_dec.emb_gru.out.gl", "erb_dec.emb_gru.linear_out.0.weight")
    g["g_df_in"] = gl("df_dec.df_gru.in.gl", "df_dec.df_gru.linear_in.0.weight")
    g["g_df_skip"] = gl("df_dec.df_skip.gl", "df_dec.df_skip.weight") if "df_dec.df_skip.weight" in sd else 0
    g["g_df_out"] = gl("df_dec.df_out.gl", "df_dec.df_out.0.weight")

    # --- GRUs
    def gru(dst: str, src: str) -> int:
        n = gru_layers(sd, src)
        for l in range(n):
            out[f"{dst}.l{l}.w_ih_t"] = f32(_np(sd[f"{src}.weight_ih_l{l}"]).T)
            out[f"{dst}.l{l}.w_ih"] = f32(_np(sd[f"{src}.weight_ih_l{l}"]))  # [3H][I]: B operand of the tcgen05 GEMM
            hi, lo = bf16_planes(_np(sd[f"{src}.weight_ih_l{l}"]))              # [3H][I] BF16 hi / lo, two per float
            out[f"{dst}.l{l}.w_ih_hi"], out[f"{dst}.l{l}.w_ih_lo"] = hi, lo
            out[f"{dst}.l{l}.w_hh"] = f32(_np(sd[f"{src}.weight_hh_l{l}"]))
            out[f"{dst}.l{l}.b_ih"] = f32(_np(sd[f"{src}.bias_ih_l{l}"]))
            out[f"{dst}.l{l}.b_hh"] = f32(_np(sd[f"{src}.bias_hh_l{l}"]))
        return n

    n_enc = gru("enc.emb_gru", "enc.emb_gru.gru")
    n_erb = gru("erb_dec.emb_gru", "erb_dec.emb_gru.gru")
    n_df = gru("df_dec.df_gru", "df_dec.df_gru.gru")
    out["enc.lsnr.w"] = f32(_np(sd["enc.lsnr_fc.0.weight"]).reshape(-1))
    out["enc.lsnr.b"] = f32(_np(sd["enc.lsnr_fc.0.bias"]).reshape(-1))
    if "df_dec.df_fc_a.0.weight" in sd:
        out["df_dec.df_fc_a.w"] = f32(_np(sd["df_dec.df_fc_a.0.weight"]).reshape(-1))
        out["df_dec.df_fc_a.b"] = f32(_np(sd["df_dec.df_fc_a.0.bias"]).reshape(-1))

    derived = dict(
        model_kind=2 if cfg.model == "deepfilternet2" else 3,
        nb_erb=cfg.nb_erb, nb_df=cfg.nb_df, df_order=cfg.df_order, df_lookahead=cfg.df_lookahead,
        conv_lookahead=cfg.conv_lookahead, conv_ch=C, conv_kt=conv_kt, inp_kt=inp_kt,
        emb_hidden=int(sd["enc.emb_gru.gru.weight_hh_l0"].shape[1]),
        df_hidden=int(sd["df_dec.df_gru.gru.weight_hh_l0"].shape[1]),
        enc_gru_layers=n_enc, erb_gru_layers=n_erb, df_gru_layers=n_df,
        df_pathway_kt=pathway_kt, enc_concat=int(cfg.enc_concat), **g,
        lsnr_scale=float(cfg.lsnr_max - cfg.lsnr_min), lsnr_offset=float(cfg.lsnr_min),
    )
    return out, derived


def random_state_dict(cfg: ModelConfig, seed: int = 0) -> Dict[str, torch.Tensor]:
    """Random-init weights with the shipped architecture's tensor names / shapes (for benchmarks
    and parity tests that must not depend on a checkpoint).  BatchNorm statistics are randomised
    too so that folding is exercised."""
    g = torch.Generator().manual_seed(seed)
    C, E, Fd = cfg.conv_ch, cfg.nb_erb, cfg.nb_df
    kt, kti = cfg.conv_kernel[0], cfg.conv_kernel_inp[0]
    H, Hd = cfg.emb_hidden_dim, cfg.df_hidden_dim
    sd: Dict[str, torch.Tensor] = {}

    def rnd(*shape, scale=1.0):
        return (torch.rand(*shape, generator=g) * 2 - 1) * scale

    def bn(p, n):
        sd[p + ".weight"] = 1.0 + rnd(n, scale=0.3)
        sd[p + ".bias"] = rnd(n, scale=0.2)
        sd[p + ".running_mean"] = rnd(n, scale=0.2)
        sd[p + ".running_var"] = 0.5 + torch.rand(n, generator=g)
        sd[p + ".num_batches_tracked"] = torch.tensor(1)

    def conv_seq(p, shapes, n_bn, pad):
        i = 1 if pad else 0
        for s in shapes:
            fan = s[1] * s[2] * s[3]
            sd[f"{p}.{i}.weight"] = rnd(*s, scale=(1.5 / fan) ** 0.5 * 1.7)
            i += 1
        bn(f"{p}.{i}", n_bn)

    conv_seq("enc.erb_conv0", [(C, 1, kti, 3)], C, kti > 1)
    for n in ("enc.erb_conv1", "enc.erb_conv2", "enc.erb_conv3", "enc.df_conv1", "erb_dec.convt3"):
        conv_seq(n, [(C, 1, kt, 3), (C, C, 1, 1)], C, kt > 1)
    conv_seq("enc.df_conv0", [(C, 1, kti, 3), (C, C, 1, 1)], C, kti > 1)
    for n in ("erb_dec.convt2", "erb_dec.convt1"):
        conv_seq(n, [(C, 1, 1, 3), (C, C, 1, 1)], C, False)
    for n in (3, 2, 1, 0):
        conv_seq(f"erb_dec.conv{n}p", [(C, 1, 1, 1)], C, False)
    conv_seq("erb_dec.conv0_out", [(1, C, kt, 3)], 1, kt > 1)
    ktp = cfg.df_pathway_kernel_size_t
    conv_seq("df_dec.df_convp", [(2 * cfg.df_order, C // 2, ktp, 1), (2 * cfg.df_order, 2 * cfg.df_order, 1, 1)],
             2 * cfg.df_order, ktp > 1)

    def glw(name, i, h, groups):
        sd[name] = rnd(groups, i // groups, h // groups, scale=(3.0 / (i // groups)) ** 0.5)

    def gruw(p, i, h, layers):
        for l in range(layers):
            k = (1.0 / h) ** 0.5
            sd[f"{p}.weight_ih_l{l}"] = rnd(3 * h, i if l == 0 else h, scale=k)
            sd[f"{p}.weight_hh_l{l}"] = rnd(3 * h, h, scale=k)
            sd[f"{p}.bias_ih_l{l}"] = rnd(3 * h, scale=k)
            sd[f"{p}.bias_hh_l{l}"] = rnd(3 * h, scale=k)

    emb_dim = C * E // 4
    if cfg.model == "deepfilternet3":
        glw("enc.df_fc_emb.0.weight", C * Fd // 2, emb_dim, cfg.enc_lin_groups)
        glw("enc.emb_gru.linear_in.0.weight", emb_dim, H, cfg.lin_groups)
        gruw("enc.emb_gru.gru", H, H, 1)
        glw("enc.emb_gru.linear_out.0.weight", H, emb_dim, cfg.lin_groups)
        sd["enc.lsnr_fc.0.weight"] = rnd(1, emb_dim, scale=0.05)
        glw("erb_dec.emb_gru.linear_in.0.weight", emb_dim, H, cfg.lin_groups)
        gruw("erb_dec.emb_gru.gru", H, H, cfg.emb_num_layers - 1)
        glw("erb_dec.emb_gru.linear_out.0.weight", H, emb_dim, cfg.lin_groups)
        glw("df_dec.df_gru.linear_in.0.weight", emb_dim, Hd, 8)  # SqueezedGRU_S default groups
        gruw("df_dec.df_gru.gru", Hd, Hd, cfg.df_num_layers)
        if cfg.df_gru_skip == "groupedlinear":
            glw("df_dec.df_skip.weight", emb_dim, Hd, cfg.lin_groups)
        glw("df_dec.df_out.0.weight", Hd, Fd * cfg.df_order * 2, cfg.lin_groups)
    else:
        G = cfg.lin_groups
        glw("enc.df_fc_emb.0.weight", C * Fd // 2, emb_dim, G)
        glw("enc.emb_gru.linear_in.0.weight", emb_dim * (2 if cfg.enc_concat else 1), H, G)
        gruw("enc.emb_gru.gru", H, H, 1)
        sd["enc.lsnr_fc.0.weight"] = rnd(1, H, scale=0.05)
        glw("erb_dec.emb_gru.linear_in.0.weight", H, H, G)
        gruw("erb_dec.emb_gru.gru", H, H, cfg.emb_num_layers - 1)
        glw("erb_dec.emb_gru.linear_out.0.weight", H, emb_dim, G)
        glw("df_dec.df_gru.linear_in.0.weight", H, Hd, 8)
        gruw("df_dec.df_gru.gru", Hd, Hd, cfg.df_num_layers)
        glw("df_dec.df_out.0.weight", Hd, Fd * cfg.df_order * 2, G)
    sd["enc.lsnr_fc.0.bias"] = rnd(1, scale=0.1)
    sd["df_dec.df_fc_a.0.weight"] = rnd(1, Hd, scale=0.05)
    sd["df_dec.df_fc_a.0.bias"] = rnd(1, scale=0.1)
    return sd
