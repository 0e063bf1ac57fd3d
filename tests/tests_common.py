"""Shared helpers of the test-suite and bench.py: synthetic noisy streams."""
import math

import torch


def synth_audio(B: int, T: int, seed: int = 1234, device: str = "cpu", sr: int = 48000) -> torch.Tensor:
    """Stream b: harmonic source (f0 ~ U[90,250] Hz, 20 harmonics with 1/h amplitudes, 4 Hz raised
    cosine envelope, RMS 0.05) + white Gaussian noise at 0 dB SNR, clipped to [-1, 1]; float32."""
    out = torch.empty((B, T), dtype=torch.float32, device=device)
    t = torch.arange(T, dtype=torch.float32, device=device) / sr
    for b in range(B):
        g = torch.Generator(device=device).manual_seed(seed + b)
        f0 = 90.0 + 160.0 * torch.rand(1, generator=g, device=device).item()
        ph = torch.rand(20, generator=g, device=device) * 2 * math.pi
        s = torch.zeros(T, dtype=torch.float32, device=device)
        for h in range(1, 21):
            if f0 * h < sr / 2:
                s += torch.sin(2 * math.pi * f0 * h * t + ph[h - 1]) / h
        env = 0.5 * (1 - torch.cos(2 * math.pi * 4.0 * t + 2 * math.pi * torch.rand(1, generator=g, device=device).item()))
        s = s * env
        s = s * (0.05 / s.pow(2).mean().sqrt().clamp_min(1e-9))
        n = torch.randn(T, generator=g, device=device) * 0.05
        out[b] = (s + n).clamp(-1, 1)
    return out
