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


# ---- transmitter (generator/ChirpGenerator.ipynb cells 1-3) -----------------------------------------
TX_T = 0.0262          # ChirpGenerator.ipynb cell 1: TIME_FRAME of the transmitter


def orth_symbol(tau, kind, f0=F0, f1=F1, T=TX_T):
    """chirp_orth law (simulation/signal.py:45-53) in continuous time; kind in 'H' (up), 'L' (down), 'G' (gap)."""
    if kind == "G":
        return np.zeros_like(tau)
    k = (f1 - f0) / T
    f = f0 + k * tau / 2.0 if kind == "H" else f1 - k * tau / 2.0
    arg = 2.0 * np.pi * f * tau - np.pi / 2.0
    return np.cos(arg) + np.sin(arg)


def frame_symbols(message, lead_in=40, guard=12):
    """G*lead_in + 7xH preamble + 1xL delimiter + bits MSB-first (H=1, L=0) + guard x G
    (ChirpGenerator.ipynb cell 2; the long lead-in covers the 24-frame mag_stat warm-up)."""
    syms = ["G"] * lead_in + ["H"] * 7 + ["L"]
    for byte in message:
        for b in range(7, -1, -1):
            syms.append("H" if (byte >> b) & 1 else "L")
    return syms + ["G"] * guard


def make_stream(message, snr_db=26.0, amp=2.0e4, start_offset=0, seed=11, lead_in=40, guard=12, nframes=None,
                fs=FS, n=N, tx_T=TX_T):
    """One receiver stream: the transmitter's frame rendered at the receiver's fs (symbols last
    TX_T = 26.2 ms = 2046.875 samples, so the lock drifts and resync has work to do), delayed by
    start_offset samples, plus Gaussian noise; int32 multiples of 256."""
    syms = frame_symbols(message, lead_in, guard)
    total = nframes * n if nframes else int(np.ceil((len(syms) * tx_T * fs + start_offset) / n)) * n
    t = (np.arange(total) - start_offset) / fs
    k = np.floor(t / tx_T).astype(np.int64)
    tau = t - k * tx_T
    x = np.zeros(total)
    valid = (k >= 0) & (k < len(syms))
    kinds = np.array([{"G": 0, "H": 1, "L": 2}[s] for s in syms])
    kk = np.where(valid, k, 0)
    kind = np.where(valid, kinds[kk], 0)
    x = np.where(kind == 1, orth_symbol(tau, "H", T=tx_T), np.where(kind == 2, orth_symbol(tau, "L", T=tx_T), 0.0)) * amp
    sig_pow = amp * amp            # mean of (cos+sin)^2 = 1
    sigma = np.sqrt(sig_pow / (10.0 ** (snr_db / 10.0)))
    x = x + np.random.default_rng(seed).standard_normal(total) * sigma
    return (np.rint(x).astype(np.int64) * 256).astype(np.int32).reshape(-1, n)


# ---- I/Q path input (simulation/IQ_modulation.ipynb cell 4: chirp_x_carrier) -----------------------
def make_iq_stream(nframes, carrier=18000.0, bw=3000.0, snr_db=10.0, amp=2.0e4, seed_bits=5, seed_noise=6, fs=FS, n=N):
    """Frames of A*cos(2*pi*(fc - fb(t))*t), fb the baseband up/down chirp -bw/2..+bw/2 over one frame
    (bit 1 = up), plus Gaussian noise, int32 x256.  -> (pcm [nframes, n], bits)."""
    t = np.arange(n) / fs
    T = n / fs
    k = bw / T
    waves = []
    for updown in ("up", "down"):
        f = -bw / 2 + k * t / 2.0 if updown == "up" else bw / 2 - k * t / 2.0
        waves.append(np.cos(2.0 * np.pi * (carrier - f) * t))
    bits = np.random.default_rng(seed_bits).integers(0, 2, nframes, dtype=np.uint8)
    sigma = np.sqrt(0.5 * amp * amp / (10.0 ** (snr_db / 10.0)))
    x = np.where(bits[:, None] == 1, waves[0][None, :], waves[1][None, :]) * amp
    x = x + np.random.default_rng(seed_noise).standard_normal((nframes, n)) * sigma
    return (np.rint(x).astype(np.int64) * 256).astype(np.int32), bits


# ---- inputs of the two earlier detectors --------------------------------------------------------------
def make_onoff_stream(bits_bytes, idle=4, amp=3.0e4, snr_db=20.0, seed=21, fs=FS, n=N, f_lo=17000.0, f_hi=19000.0):
    """experiments/chirp framing (Inc/main.h:95-98): a byte = 3 HIGH frames (sync), then 8 bits x 2 frames each
    (HIGH frame pair = 1, silent pair = 0), one frame of slack; HIGH = a linear chirp across the counted band."""
    t = np.arange(n) / fs
    T = n / fs
    chirp = np.sin(2.0 * np.pi * (f_lo * t + 0.5 * (f_hi - f_lo) / T * t * t))
    levels = [0] * idle
    for byte in bits_bytes:
        levels += [1, 1, 1]
        for b in range(7, -1, -1):
            levels += [1, 1] if (byte >> b) & 1 else [0, 0]
        levels += [0] * (idle + 1)
    sigma = np.sqrt(0.5 * amp * amp / (10.0 ** (snr_db / 10.0)))
    rng = np.random.default_rng(seed)
    x = np.stack([chirp * amp * lv for lv in levels]) + rng.standard_normal((len(levels), n)) * sigma
    return (np.rint(x).astype(np.int64) * 256).astype(np.int32), np.array(levels, np.int8)


def make_fsk_stream(message, repeat=4, amp=3.0e4, snr_db=20.0, seed=22, n=N, sof=340, eof=344, hex0=348, step=4, gap=2):
    """experiments/ultracom framing: start tone, two hex-digit tones per byte (MSB nibble first), end tone; every
    tone lasts `repeat` frames (parser() wants a code three frames in a row at TQ_N = 2), `gap` silent frames
    between tones so repeated digits are seen as new codes."""
    bins = [sof]
    for byte in message:
        bins += [hex0 + step * (byte >> 4), hex0 + step * (byte & 15)]
    bins.append(eof)
    t = np.arange(n)
    frames, codes = [], []
    for b in bins:
        frames += [np.sin(2.0 * np.pi * b * t / n) * amp] * repeat + [np.zeros(n)] * gap
        codes += [b] * repeat + [0] * gap
    sigma = np.sqrt(0.5 * amp * amp / (10.0 ** (snr_db / 10.0)))
    x = np.stack(frames) + np.random.default_rng(seed).standard_normal((len(frames), n)) * sigma
    return (np.rint(x).astype(np.int64) * 256).astype(np.int32), np.array(codes)
