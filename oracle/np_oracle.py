"""numpy (float64) oracle — TEST INFRASTRUCTURE ONLY (see oracle/ref_dsp.h for the rules).

Restates the reference's Python simulation library (simulation/chirp.py, dsp.py, signal.py) and the
notebook chains built on it, plus float64 versions of the firmware chains used to bound the fp32
C oracle / CUDA results at the north-star tolerance (1e-4 relative on magnitudes).

`load_reference_simulation()` imports the real reference modules by file path with the plotting
modules stubbed; it only works where /root/reference exists (this container) and is used by
tests/golden/make_golden.py to generate the committed golden vectors.  Nothing under tests -m gpu,
smoke() or bench.py calls it.
"""
import importlib.util
import os
import sys
import types

import numpy as np

REFERENCE_ROOT = "/root/reference"


# ---- simulation/chirp.py:29-37 == dsp.py:117-125 == signal.py:28-36 ---------------------------
def chirp(f0, f1, fs, T, amp=1.0, updown="up", phase=-np.pi / 2.0):
    """Complex chirp.  NOTE linspace includes the endpoint: dt = T/(N-1), N = int(T*fs)."""
    t = np.linspace(0, T, int(T * fs))
    k = float(f1 - f0) / float(T)
    f = f0 + k * t / 2.0 if updown == "up" else f1 - k * t / 2.0
    return np.exp(1j * ((2.0 * np.pi * f * t) + phase)) * amp


def chirp_cos(*a, **kw):          # chirp.py:40-48
    return np.real(chirp(*a, **kw))


def chirp_sin(*a, **kw):          # chirp.py:51-59
    return np.imag(chirp(*a, **kw))


def chirp_orth(f0, f1, fs, T, amp=1.0, updown="up", phase=-np.pi / 2.0):   # signal.py:45-53
    c = chirp(f0, f1, fs, T, 1.0, updown, phase)
    return (np.real(c) + np.imag(c)) * amp


def white_noise(n, amp, rng):     # chirp.py:62-65 / dsp.py:136-141: UNIFORM(-A, A) + 1j*UNIFORM(-A, A)
    a = rng.random(n) * 2 * amp - amp
    b = rng.random(n) * 2 * amp - amp
    return a + 1j * b


def time_shift(x, shift_rate):    # chirp.py:126-129 / signal.py:134-137
    t = int(len(x) * shift_rate)
    return np.append(x[t:], x[:t])


def add_delay(x, delay_rate):     # chirp.py:118-123 / signal.py:128-132
    n = len(x)
    la = int(n * delay_rate)
    return np.append(np.append(np.zeros(la), x), np.zeros(2 * n - (n + la)))


def buffer(fs, length, wave):     # dsp.py:83-87
    buf = np.zeros(int(fs * length))
    buf[:len(wave)] = wave
    return buf


def fft_peak_freqs(wave, fs, thres=0.95):
    """plot_fft's printed result (chirp.py:74-90): frequencies of the local maxima above
    min + thres*(max-min) — peakutils.indexes(a, thres) restated (min_dist=1)."""
    y = np.fft.fftshift(np.fft.fft(wave))
    freq = np.fft.fftshift(np.fft.fftfreq(len(y), 1 / fs))
    a = np.abs(y)
    th = thres * (a.max() - a.min()) + a.min()
    d = np.diff(a)
    # peakutils: first-order difference sign change (plateaus resolved to the left edge)
    peaks = np.where((np.hstack([d, 0.0]) < 0.0) & (np.hstack([0.0, d]) > 0.0) & (a > th))[0]
    return freq[peaks]


def dechirp_spectrum(rx, ref):    # ChirpSimulation.ipynb cell 31: fft((c + noise) * conj(c))
    return np.fft.fft(rx * ref)


def compress(rx, ref):            # ChirpSimulation.ipynb cells 36, 38: ifft(fft(rx) * fft(ref))
    return np.fft.ifft(np.fft.fft(rx) * np.fft.fft(ref))


# ---- float64 versions of the firmware chains (same tables as the fp32 chain, exact arithmetic) --
def receiver_mags_f64(frame, chirp_tab, hann):
    """receiver/Src/main.c:163-180 in float64 on the given (fp32-valued) tables: packed-bin mags."""
    x = np.asarray(frame, np.float64) * np.asarray(chirp_tab, np.float64) * np.asarray(hann, np.float64)
    X = np.fft.rfft(x)
    mag = np.abs(X[:len(x) // 2])
    mag[0] = np.hypot(X[0].real, X[len(x) // 2].real)       # packed DC/Nyquist "bin 0"
    return mag


def pack_rfft(X, n):
    """numpy rfft (n/2+1 bins) -> CMSIS packed layout (arm_math.h:2246-2249 contract)."""
    out = np.empty(n, np.float64)
    out[0] = X[0].real
    out[1] = X[n // 2].real
    out[2::2] = X[1:n // 2].real
    out[3::2] = X[1:n // 2].imag
    return out


def unpack_rfft(p):
    n = len(p)
    X = np.empty(n // 2 + 1, np.complex128)
    X[0] = p[0]
    X[n // 2] = p[1]
    X[1:n // 2] = p[2::2] + 1j * p[3::2]
    return X


def compress_chain_f64(frame, window, H_packed, quirk=True):
    """experiments/chirp_compression_time_domain/Src/chirp.c:78-83 in float64.
    quirk=True reproduces the packed (X0, X_{N/2}) complex product of arm_cmplx_mult_cmplx_f32."""
    n = len(frame)
    x = np.asarray(frame, np.float64) * np.asarray(window, np.float64)
    P = pack_rfft(np.fft.rfft(x), n)
    H = np.asarray(H_packed, np.float64)
    a = P[0::2] + 1j * P[1::2]
    b = H[0::2] + 1j * H[1::2]
    prod = a * b
    if not quirk:
        prod[0] = P[0] * H[0] + 1j * (P[1] * H[1])
    Q = np.empty(n, np.float64)
    Q[0::2] = prod.real
    Q[1::2] = prod.imag
    return np.fft.irfft(unpack_rfft(Q), n)


# ---- the real reference modules (this container only) ------------------------------------------
def load_reference_simulation(name):
    """Import /root/reference/simulation/<name>.py by path with plotting modules stubbed.
    Never put simulation/ on sys.path: its signal.py shadows the stdlib module."""
    for m in ("matplotlib", "matplotlib.pyplot", "peakutils", "IPython", "IPython.display"):
        if m not in sys.modules:
            stub = types.ModuleType(m)
            stub.rcParams = {}
            stub.display = lambda *a, **k: None
            stub.Audio = lambda *a, **k: None
            sys.modules[m] = stub
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    path = os.path.join(REFERENCE_ROOT, "simulation", name + ".py")
    spec = importlib.util.spec_from_file_location("refsim_" + name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod
