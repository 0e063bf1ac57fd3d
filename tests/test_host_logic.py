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
    p = tmp_path / "config.ini"
    p.write_text("[train]\nmodel = deepfilternet\n")
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
