"""CPU: host-side logic -- config parsing, weight packing, checkpoint selection, the FFT index
algebra (host emulation of the warp kernel), and that libdfb200.so exports the whole C ABI."""
import ctypes
import os
import re
import subprocess

import numpy as np
import pytest
import torch

from conftest import ROOT

from deepfilternet_b200 import _lib
from deepfilternet_b200.config import ModelConfig, load_config
from deepfilternet_b200.model import find_checkpoint
from deepfilternet_b200.weights import pack_state_dict, random_state_dict


def dfn3_cfg(**kw):
    d = dict(model="deepfilternet3", conv_ch=64, conv_lookahead=2, df_lookahead=2, emb_num_layers=3, df_num_layers=2,
             lin_groups=16, enc_lin_groups=32, df_gru_skip="groupedlinear", df_pathway_kernel_size_t=5)
    d.update(kw)
    return ModelConfig(**d)


def test_capi_exports_every_declared_symbol():
    """Every function declared in include/dfb200.h is exported by the built library (no compute)."""
    hdr = open(os.path.join(ROOT, "include", "dfb200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    names = set(re.findall(r"\b(dfb_[a-z0-9_]+)\s*\(", hdr))
    assert len(names) >= 30
    if not os.path.exists(_lib.SO_PATH):
        _lib.build()
    L = ctypes.CDLL(_lib.SO_PATH)
    for n in sorted(names):
        assert hasattr(L, n), f"{n} declared in dfb200.h but not exported"
    assert set(_lib.SIGNATURES) == names
    assert b"sm_100a" in ctypes.cast(L.dfb_version, ctypes.CFUNCTYPE(ctypes.c_char_p))()


def test_product_fails_loudly_without_gpu():
    """No CPU fallback: without a usable device every entry point reports DFB_ERR_CUDA."""
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from deepfilternet_b200 import libdf
    with pytest.raises(RuntimeError, match="no CUDA device|CUDA"):
        libdf.DF(48000, 960, 480, 32, 2)
    with pytest.raises(RuntimeError):
        libdf.erb(np.zeros((2, 481), np.complex64), np.array([481], np.uint64))


def test_erb_widths_host_entry(golden_dir):
    import json
    g = json.load(open(os.path.join(golden_dir, "erb_widths.json")))
    out = (ctypes.c_int64 * 32)()
    assert _lib.lib().dfb_erb_widths(48000, 960, 32, 2, out) == 0
    assert list(out) == g["from_checkpoint_erb_fb"]["DeepFilterNet3"]
    import libdf_oracle as LO
    for args in [(24000, 192, 24, 1), (48000, 960, 32, 1), (16000, 512, 24, 2), (44100, 1024, 40, 3)]:
        o = (ctypes.c_int64 * args[2])()
        assert _lib.lib().dfb_erb_widths(*args, o) == 0
        assert list(o) == LO.erb_widths(*args).tolist()


def test_config_parsing(model_dir, tmp_path):
    c3 = load_config(os.path.join(model_dir, "DeepFilterNet3", "config.ini"), env={})
    assert (c3.model, c3.conv_lookahead, c3.df_lookahead, c3.emb_hidden_dim, c3.lin_groups, c3.enc_lin_groups) == \
        ("deepfilternet3", 2, 2, 256, 16, 32)
    assert c3.norm_alpha == 0.99
    c2 = load_config(os.path.join(model_dir, "DeepFilterNet2", "config.ini"), env={})
    assert c2.model == "deepfilternet2" and c2.enc_concat and c2.lin_groups == 8 and c2.df_order == 5 and c2.df_lookahead == 2
    cl = load_config(os.path.join(model_dir, "DeepFilterNet3_ll", "config.ini"), env={})
    assert (cl.conv_lookahead, cl.df_lookahead, tuple(cl.conv_kernel), cl.emb_hidden_dim, cl.df_num_layers) == (0, 0, (2, 3), 512, 3)
    # df/config.py:119-122: environment variables named like the option win over the ini file
    ce = load_config(os.path.join(model_dir, "DeepFilterNet3", "config.ini"), env={"DF_ORDER": "3"})
    assert ce.df_order == 3
    c1 = load_config(os.path.join(model_dir, "DeepFilterNet", "config.ini"), env={})
    assert (c1.model, c1.conv_lookahead, c1.df_lookahead, c1.emb_hidden_dim, c1.gru_groups, c1.lin_groups, tuple(c1.conv_kernel_inp)) == \
        ("deepfilternet", 2, 1, 512, 8, 8, (2, 3))
    p = tmp_path / "config.ini"
    p.write_text("[train]\nmodel = deepfilternet\n")   # v1 with the code defaults (conv_ch 16, one decoder time tap, ...): not built
    with pytest.raises(NotImplementedError):
        load_config(str(p), env={})


@pytest.mark.parametrize("cfg", [dfn3_cfg(), dfn3_cfg(conv_lookahead=0, df_lookahead=0, conv_kernel=(2, 3), emb_hidden_dim=512,
                                                       df_hidden_dim=512, df_num_layers=3, enc_lin_groups=16),
                                 ModelConfig(model="deepfilternet2", conv_ch=64, conv_lookahead=2, df_lookahead=2, emb_num_layers=3,
                                             df_num_layers=2, lin_groups=8, enc_lin_groups=8, enc_concat=True,
                                             df_pathway_kernel_size_t=5)])
def test_pack_random_weights_and_oracle_forward(cfg):
    import dfnet_oracle as O
    sd = random_state_dict(cfg, seed=0)
    packed, derived = pack_state_dict(sd, cfg)
    assert packed["enc.erb_conv1.pw"].shape == (64, 64) and packed["enc.erb_conv0.w"].shape == (3, 3, 64)
    assert derived["conv_kt"] == cfg.conv_kernel[0] and derived["df_pathway_kt"] == 5
    assert packed["enc.emb_gru.l0.w_ih_t"].shape == (cfg.emb_hidden_dim, 3 * cfg.emb_hidden_dim)
    assert all(a.dtype == np.float32 and a.flags["C_CONTIGUOUS"] for a in packed.values())
    # BN folding: y = conv(x) * s + b  ==  conv'(x) + b'
    x = torch.randn(1, 64, 3, 8)
    ref = O.conv_norm_act(x, sd, "enc.erb_conv3", act="relu")
    kt = cfg.conv_kernel[0]
    xp = torch.nn.functional.pad(x, (1, 1, kt - 1, 0))
    dw = torch.from_numpy(packed["enc.erb_conv3.dw"])  # [kt][3][C]
    a = sum(xp[:, :, dt:dt + 3, df:df + 8] * dw[dt, df].view(1, 64, 1, 1) for dt in range(kt) for df in range(3))
    y = torch.einsum("bctf,cn->bntf", a, torch.from_numpy(packed["enc.erb_conv3.pw"])) + torch.from_numpy(packed["enc.erb_conv3.b"]).view(1, 64, 1, 1)
    assert torch.allclose(torch.relu(y), ref, atol=2e-5)
    # the oracle runs end to end on these weights
    audio = torch.randn(1, 4800) * 0.05
    out = O.enhance(sd, cfg.as_dict(), audio)
    assert out.shape == audio.shape and torch.isfinite(out).all()


def test_checkpoint_selection(tmp_path):
    d = tmp_path / "checkpoints"
    d.mkdir()
    for n in ("model_10.ckpt", "model_96.ckpt.best", "model_120.ckpt"):
        (d / n).write_bytes(b"")
    assert find_checkpoint(str(d), "best")[1] == 96
    assert find_checkpoint(str(d), "latest")[1] == 120
    assert find_checkpoint(str(d), 10)[1] == 10
    assert find_checkpoint(str(tmp_path), "best") == (None, None)


def test_fft_index_algebra_on_host(tmp_path):
    """Host emulation of the one-warp 480-point FFT + real split/merge used by the CUDA kernels."""
    exe = tmp_path / "fft_host_test"
    subprocess.check_call(["nvcc", "-std=c++17", "-O1", "-Wno-deprecated-gpu-targets", "-o", str(exe),
                           os.path.join(ROOT, "tests", "host", "fft_host_test.cu")], stderr=subprocess.DEVNULL)
    out = subprocess.run([str(exe)], capture_output=True, text=True)
    assert out.returncode == 0 and "OK" in out.stdout, out.stdout


def test_umma_sw128_image_layout():
    """The pre-swizzled BF16 hi | lo image of the 1x1 conv weights (the tcgen05 B operand the conv kernel fetches with
    one bulk copy): un-swizzling chunk j ^ (n & 7) of row n must give back hi + lo ~= w to BF16x2 accuracy."""
    import torch
    from deepfilternet_b200.weights import bf16_planes, umma_sw128_image
    rng = np.random.default_rng(0)
    w = rng.standard_normal((64, 64)).astype(np.float32)
    img = umma_sw128_image(w)
    assert img.dtype == np.float32 and img.size == 2 * 64 * 32  # two planes of 64 rows x 128 bytes
    planes = img.view(np.int16).reshape(2, 64, 8, 8)             # [plane][row][16-byte chunk][8 bf16]
    rec = np.zeros((2, 64, 64), dtype=np.float32)
    for n in range(64):
        for j in range(8):
            chunk = planes[:, n, j ^ (n & 7)]                    # where chunk j of row n was stored
            rec[:, n, 8 * j:8 * j + 8] = torch.from_numpy(chunk.copy()).view(torch.bfloat16).to(torch.float32).numpy()
    hi, lo = bf16_planes(w)
    as_f32 = lambda p: torch.from_numpy(p.view(np.int16).copy()).view(torch.bfloat16).to(torch.float32).numpy().reshape(64, 64)
    assert np.array_equal(rec[0], as_f32(hi)) and np.array_equal(rec[1], as_f32(lo))
    assert np.abs(rec[0] + rec[1] - w).max() <= 2.0 ** -16 * np.abs(w).max()


def test_bench_bookkeeping():
    """bench.py's clock parser and kernel table (no GPU): nvidia-smi csv lines -> median clock / throttle reasons
    inside the timed window; every kernel name the library's profiler can emit has a roofline model entry."""
    import importlib.util
    import re
    import time
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("bench", os.path.join(root, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    now = time.time()
    mk = lambda ts, sm, cap: (ts, f"2026/01/01 00:00:00.000, 0, {sm}, 1965, 600.0, 0x0, Not Active, Not Active, Not Active, {cap}")
    lines = [mk(now - 2.0, 1200, "Not Active"), mk(now - 0.2, 1950, "Active"), mk(now - 0.1, 1965, "Not Active")]
    r = bench.ClockSampler.parse(lines, now - 0.3, now)
    assert r["samples"] == 2 and r["sm_mhz"] == 1965.0 and r["reasons"] == ["sw_power_cap"] and r["scope"] == "timed regions"
    r = bench.ClockSampler.parse(lines, now + 1, now + 2)
    assert r["samples"] == 3 and r["scope"].startswith("warm-up")
    names = set()
    csrc = os.path.join(root, "deepfilternet_b200", "csrc")
    for f in os.listdir(csrc):
        if f.endswith(".cu"):
            names |= set(re.findall(r'DFB_PROF\("([^"]+)"', open(os.path.join(csrc, f)).read()))
    # (k_gather_sum / k_convp_v1 only run for DeepFilterNet v1, which is not a BASELINE config)
    default_path = {n for n in names if not n.startswith(("k_gru", "k_dwpw", "k_mask_out", "k_to_planes", "k_gather_sum", "k_convp_v1"))
                    or n in ("k_gru_tc", "k_dwpw_bx")}
    for model_name in ("DeepFilterNet3", "DeepFilterNet2", "DeepFilterNet3_ll"):
        cfg = bench.model_config(model_name)
        from deepfilternet_b200.weights import pack_state_dict, random_state_dict
        _, derived = pack_state_dict(random_state_dict(cfg, seed=0), cfg)
        km = bench.kernel_model(cfg, derived)
        assert default_path <= set(km), default_path - set(km)
        assert all(v[1] > 0 for v in km.values())
    # DFN3: the GRU figure is SURVEY 8d's 1 966 080 MAC split evenly into recurrence and projection
    cfg = bench.model_config("DeepFilterNet3")
    _, derived = pack_state_dict(random_state_dict(cfg, seed=0), cfg)
    km = bench.kernel_model(cfg, derived)
    assert km["k_gru_tc"][1] == 2 * 5 * 256 * 768 and km["k_analysis"][1] == 1920 + 3848 + 128
    assert km["k_apply_synthesis"][1] == 3848 + 128 + 3840 + 1920 and km["k_dwpw_bx"][1] == 256 * 352
    assert set(bench.CONFIGS) == {2, 3, 4, 5, 6} and bench.CONFIGS[2] == ("DeepFilterNet3", 128, 10, 1)
    # DeepFilterNet v1 (cfg 6, not a BASELINE config): the grouped recurrences count 1 / G of the dense product
    cfg = bench.model_config("DeepFilterNet")
    _, derived = pack_state_dict(random_state_dict(cfg, seed=0), cfg)
    km = bench.kernel_model(cfg, derived)
    assert km["k_gru_tc512"][1] == 2 * 3 * 512 * 512 * 5 // 8 and km["k_gather_sum"][1] > 0 and km["k_convp_v1"][1] > 0
    # per-GPU report: one row per rank, NaN -> null, unknown kernel -> null, malformed rows never raise
    import json
    rep = bench.per_gpu_report([[10.0, 12.5, 0.02, 0.8, 1.0, 0.3], [11.0, 13.0, float("nan"), float("nan"), -1.0, 0.4]], ["a", "b"], 128, 10)
    assert rep[0]["value"] == 128000.0 and rep[0]["roofline_kernel"] == "b" and rep[1]["roofline_kernel"] is None
    assert rep[1]["roofline_frac"] is None and json.loads(json.dumps(rep))[1]["rank"] == 1
    assert "error" in bench.per_gpu_report([[1.0]], ["a"], 1, 1)[0]




def test_io_wav_roundtrip_and_resample_taps(tmp_path):
    """df.io mirror, host side: the RIFF reader / writer (PCM16 with the reference's 1 << 15 scaling, float32, 24-bit) and
    the resampler taps against torchaudio's own `_get_sinc_resample_kernel` (the convolution itself runs on the GPU)."""
    import math
    import struct
    from deepfilternet_b200 import io as dio
    rng = np.random.default_rng(0)
    x = (rng.standard_normal((2, 1000)) * 0.1).astype(np.float32)
    p = dio.save_audio(str(tmp_path / "a.wav"), torch.from_numpy(x), 48000, suffix="enh")
    assert p.endswith("a_enh.wav")
    y, meta = dio.load_audio(p)
    assert (meta.sample_rate, meta.num_frames, meta.num_channels, meta.bits_per_sample, meta.encoding) == (48000, 1000, 2, 16, "PCM_S")
    assert np.array_equal(y.numpy(), (x * 32768).astype(np.int16).astype(np.float32) / 32768)
    p = dio.save_audio(str(tmp_path / "f.wav"), torch.from_numpy((x * 32768).astype(np.int16)), 16000, dtype=torch.float32)
    y, meta = dio.load_audio(p, frame_offset=10, num_frames=100)
    assert meta.encoding == "PCM_F" and y.shape == (2, 100)
    # 24-bit PCM, hand-made
    v = np.array([0, 1, -1, (1 << 23) - 1, -(1 << 23)], dtype=np.int32)
    raw = b"".join(struct.pack("<i", int(a))[:3] for a in v)
    hdr = struct.pack("<4sI4s4sIHHIIHH4sI", b"RIFF", 36 + len(raw), b"WAVE", b"fmt ", 16, 1, 1, 8000, 24000, 3, 24, b"data", len(raw))
    (tmp_path / "p24.wav").write_bytes(hdr + raw)
    y, meta = dio.load_audio(str(tmp_path / "p24.wav"))
    assert meta.bits_per_sample == 24 and np.allclose(y.numpy()[0], v / float(1 << 23))
    with pytest.raises(RuntimeError):
        (tmp_path / "bad.wav").write_bytes(b"not a wav file at all")
        dio.load_audio(str(tmp_path / "bad.wav"))
    try:
        from torchaudio.functional.functional import _get_sinc_resample_kernel
    except Exception:
        pytest.skip("torchaudio not importable")
    for o, n, meth in [(44100, 48000, "sinc_fast"), (48000, 16000, "kaiser_best"), (16000, 48000, "kaiser_fast"), (48000, 44100, "sinc_best")]:
        k, w, og, nw = dio.resample_kernel(o, n, **dio.get_resample_params(meth))
        kr, wr = _get_sinc_resample_kernel(o, n, math.gcd(o, n), **dio.get_resample_params(meth))
        assert w == wr and (og, nw) == (o // math.gcd(o, n), n // math.gcd(o, n))
        assert torch.equal(k, kr[:, 0])


def test_v1_packing_against_oracle_through_device_graph_emulation():
    """DeepFilterNet v1: the packed tensors (block-diagonal GRUs with folded shuffles, gather tables, reversed transposed-conv
    time taps, df_fc_out in the device coefs layout) run through a torch emulation of the DEVICE graph (tests/v1_emulation.py:
    channel-last tensors, the same kernel sequence as forward_v1) must reproduce the oracle."""
    import dfnet1_oracle as O1
    import libdf_oracle as LO
    import v1_emulation as em
    cfg = ModelConfig(model="deepfilternet", conv_lookahead=2, df_lookahead=1, conv_ch=64, conv_kernel=(2, 3), convt_kernel=(2, 3),
                      conv_kernel_inp=(2, 3), conv_k_enc=2, conv_k_dec=2, emb_hidden_dim=512, df_hidden_dim=512, emb_num_layers=3,
                      df_num_layers=2, gru_groups=8, lin_groups=8, enc_lin_groups=8, group_shuffle=True, dfop_method="real_unfold")
    sd = random_state_dict(cfg, seed=3)
    packed, d = pack_state_dict(sd, cfg)
    assert d["model_kind"] == 1 and d["enc_gru_layers"] == 3 and d["df_gru_layers"] == 2
    g = torch.Generator().manual_seed(0)
    B, T = 2, 11
    fe = torch.randn(B, 1, T, 32, generator=g) * 0.5
    fs = torch.randn(B, 1, T, 96, 2, generator=g) * 0.5
    spec = torch.randn(B, 1, T, 481, 2, generator=g)
    widths = LO.erb_widths(48000, 960, 32, 2)
    spec_e, m, lsnr, co, alpha = O1.dfnet1_forward(sd, dict(O1.DEFAULTS_DFN1), widths, spec, fe, fs)
    m2, coefs2, lsnr2, alpha2 = em.forward(packed, d, fe[:, 0], fs[:, 0])
    co_dev = co.permute(0, 1, 3, 2, 4).reshape(B, T, -1)   # [B,T,O,Fd,2] -> device layout [B,T,Fd * 2 O]
    assert float((m2 - m[:, 0]).abs().max()) < 1e-5 and float((coefs2 - co_dev).abs().max()) < 1e-5
    assert float((lsnr2 - lsnr[..., 0]).abs().max()) < 1e-4 and float((alpha2 - alpha[..., 0]).abs().max()) < 1e-5
