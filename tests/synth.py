"""Seeded synthetic receiver input, host side, for parity subsets.

Each frame = A * chirp_orth(up|down) (simulation/signal.py:45-53 law, restated in
oracle/np_oracle.py, at fs = 78125 Hz, T = N/fs so a frame is exactly N samples) + Gaussian noise,
quantised to int32 multiples of 256 like the raw DFSDM words (receiver/Src/dfsdm.c:78: 24-bit
sample in bits 31:8, no shift applied anywhere).  Bits come from rng(seed_bits), noise from
rng(seed_noise).
"""
import numpy as np

from oracle import np_oracle as NP

N = 2048
FS = 78125.0
F0, F1 = 16000.0, 19000.0


def symbol_waves(n=N, fs=FS, f0=F0, f1=F1):
    T = n / fs
    while int(T * fs) < n:          # guard against float truncation in int(T*fs)
        T = np.nextafter(T, np.inf)
    up = NP.chirp_orth(f0, f1, fs, T, 1.0, "up")
    down = NP.chirp_orth(f0, f1, fs, T, 1.0, "down")
    assert len(up) == n and len(down) == n
    return up, down


def make_frames(nframes, snr_db=-5.0, amp=2.0e4, seed_bits=2, seed_noise=3, n=N, dtype=np.int32):
    """-> (pcm [nframes, n] int32 multiples of 256 (or float32 of the same values), bits [nframes])."""
    up, down = symbol_waves(n)
    bits = np.random.default_rng(seed_bits).integers(0, 2, size=nframes, dtype=np.uint8)
    sig_pow = np.mean(up ** 2) * amp * amp
    sigma = np.sqrt(sig_pow / (10.0 ** (snr_db / 10.0)))
    rng = np.random.default_rng(seed_noise)
    out = np.empty((nframes, n), np.int32)
    chunk = 4096
    for s in range(0, nframes, chunk):
        e = min(nframes, s + chunk)
        sym = np.where(bits[s:e, None] == 1, up[None, :], down[None, :]) * amp
        x = sym + rng.standard_normal((e - s, n)) * sigma
        out[s:e] = (np.rint(x).astype(np.int64) * 256).astype(np.int32)
    if dtype == np.float32:
        return out.astype(np.float32), bits
    return out, bits
