"""Golden vectors for the optional stages of the enhancement path -- the post filter (`init_df(post_filter=True)`,
deepfilternet3.py:448-454 / modules.py:234-245) and `mask_only` (enhance.py:172-175, checkpoint.py:32) -- produced by the
REFERENCE's own modules (imported from /root/reference; build container only) on the excerpt of gen_golden.py.
Run:  python oracle/gen_golden_pf.py   ->  tests/golden/dfnet_pf.npz
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
import ref_harness as rh  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")


def main():
    d = rh.unpack_models()
    rh.import_reference()
    from df.enhance import enhance, init_df
    noisy = torch.from_numpy(rh.read_wav(os.path.join(GOLD, "assets", "noisy_snr0.wav")))
    x = torch.stack([noisy[0, 96000:120000], 0.5 * noisy[0, 130000:154000]])
    out = {"audio": x.numpy()}
    for name in ("DeepFilterNet3", "DeepFilterNet2"):
        model, st, suffix, _ = init_df(os.path.join(d, name), post_filter=True, log_file=None, log_level="ERROR")
        assert suffix.endswith("_pf")
        out[f"{name}_pf"] = enhance(model, st, x, pad=True).numpy()
        out[f"{name}_pf_atten12"] = enhance(model, st, x, pad=True, atten_lim_db=12.0).numpy()
        model, st, _, _ = init_df(os.path.join(d, name), mask_only=True, log_file=None, log_level="ERROR")
        out[f"{name}_mask_only"] = enhance(model, st, x, pad=True).numpy()
    np.savez_compressed(os.path.join(GOLD, "dfnet_pf.npz"), **out)
    print({k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
