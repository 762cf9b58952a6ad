"""Generate the committed golden fixtures from the reference (run in the build container only:
needs /root/reference).  Outputs (all under tests/golden/):

  device_triples.npz   the 24 device captures agent/{chirp_experiment,vaccum_cleaner}/*.{raw,flt,fft}
                       produced by real CMSIS-DSP on the Cortex-M4 (experiments/basic/Src/main.c:107-175):
                       raw int32 PCM, PCM x Hann (6-decimal print), |rfft|/sqrt(N) (<1 kHz forced to 1.0).
                       `consistent` flags the 23 triples whose raw/fft come from the same capture.
  refsim_vectors.npz   outputs of the reference's own Python (simulation/chirp.py, dsp.py, signal.py
                       imported by path): chirps, chirp_orth, time_shift, add_delay, and the
                       ChirpSynchronization.ipynb cell 5-11 products with the printed peak frequencies.
  fir_taps.npz         the 27 FIR taps literal at experiments/iq_modulation/Src/iq_modem.c:18.
  tx_vectors.npz       the transmitter of generator/ChirpGenerator.ipynb cells 1-3 run through the reference's own
                       simulation/signal.py: H, L symbols at 44.1 kHz and the whole int16 tone of the message "Hi"
                       (G + 7H + L + bits + 12G, astype(int16) as Signal.play() writes it to the WAV file).
  wire_excerpt.json    verbatim head / tail lines of one capture's .raw / .flt / .fft files (the PC agent's
                       CSV wire formats, agent/README.md:5-11) with the row indices they came from.
"""
import glob
import os
import re
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import np_oracle  # noqa: E402

REF = "/root/reference"
N = 2048


def load_csv(path):
    return np.loadtxt(path, delimiter=",", skiprows=1)


def device_triples():
    names, raws, flts, ffts, freqs, ok = [], [], [], [], [], []
    for d in ("chirp_experiment", "vaccum_cleaner"):
        for rawf in sorted(glob.glob(os.path.join(REF, "agent", d, "*.raw"))):
            base = rawf[:-4]
            if not (os.path.exists(base + ".flt") and os.path.exists(base + ".fft")):
                continue
            raw = load_csv(rawf)[:, 1]
            flt = load_csv(base + ".flt")[:, 1]
            fft = load_csv(base + ".fft")
            if len(raw) != N or len(flt) != N or len(fft) != N // 2:
                continue
            names.append(d + "/" + os.path.basename(base))
            raws.append(raw.astype(np.int32))
            flts.append(flt)
            freqs.append(fft[:, 0])
            ffts.append(fft[:, 1])
            # 48.1(kHz)_M2A's raw and fft come from different captures
            ok.append(os.path.basename(base) != "48.1(kHz)_M2A")
    np.savez_compressed(os.path.join(HERE, "device_triples.npz"), names=np.array(names),
                        raw=np.array(raws), flt=np.array(flts), fft_mag=np.array(ffts),
                        fft_freq=np.array(freqs), consistent=np.array(ok))
    print("device triples:", len(names), "consistent:", int(np.sum(ok)))


def refsim_vectors():
    sig = np_oracle.load_reference_simulation("signal")
    ch = np_oracle.load_reference_simulation("chirp")
    dsp = np_oracle.load_reference_simulation("dsp")
    out = {}
    # ChirpSynchronization.ipynb cell 3
    s = sig.Signal(f0=16000, f1=19000, fs=100000, T=0.0205, A=20000)
    out["sync_chirp"] = s.chirp()
    out["sync_chirp_cos"] = s.chirp_cos()
    out["sync_chirp_down"] = s.chirp(updown="down")
    out["sync_chirp_orth"] = s.chirp_orth()
    shifts = [0.0, 1.0 / 8.0, 2.0 / 8.0, 4.0 / 8.0]
    out["sync_shift_rates"] = np.array(shifts)
    for i, r in enumerate(shifts):
        out["sync_product_%d" % i] = sig.time_shift(s.chirp_cos(), r) * s.chirp()      # cell 3
    # printed outputs of cells 5, 7, 9, 11
    out["sync_peak_hz_0"] = np.array([0.0])
    out["sync_peak_hz_1"] = np.array([-390.24390244])
    out["sync_peak_hz_2"] = np.array([-731.70731707])
    out["sync_peak_hz_3"] = np.array([-1512.19512195, 1512.19512195])
    # chirp.py free functions and dsp.Chirp at the config-1 shape (N = 2048 exactly)
    out["cfg1_chirp_up"] = ch.chirp(f0=16000, f1=18000, fs=100000, T=0.02048, amp=1.0)
    out["cfg1_chirp_down"] = ch.chirp(f0=16000, f1=18000, fs=100000, T=0.02048, amp=1.0, updown="down")
    c = dsp.Chirp(fs=100000, f0=16000, f1=18000, T=0.02048, A=1.0)
    out["cfg1_dsp_chirp_up"] = c.chirp()
    out["cfg1_dsp_chirp_cos_down_p"] = c.chirp_cos(updown="down", phase=np.pi / 2)
    out["delay_0p25"] = sig.add_delay(np.arange(8.0), 0.25)
    out["shift_0p25"] = sig.time_shift(np.arange(8.0), 0.25)
    # OrthogonalChirp.ipynb cells 2, 8, 11-13 (noise-free): W x up -> 0 Hz, W x down -> 35024.39 Hz
    sr = sig.Signal(f0=16000, f1=19000, fs=100000, T=0.0205, A=20000)
    W = np.real(sr.chirp()) + np.imag(sr.chirp())
    out["orth_Ru"] = W * sr.chirp(updown="up")
    out["orth_Rd"] = W * sr.chirp(updown="down")
    out["orth_peak_hz_u"] = np.array([0.0])
    out["orth_peak_hz_d"] = np.array([35024.3902439])
    np.savez_compressed(os.path.join(HERE, "refsim_vectors.npz"), **out)
    print("refsim vectors:", len(out))


def fir_taps():
    src = open(os.path.join(REF, "experiments/iq_modulation/Src/iq_modem.c")).read()
    m = re.search(r"float32_t b\[27\] = \{([^}]*)\}", src)
    taps = np.array([float(v) for v in m.group(1).split(",")], dtype=np.float64)
    assert taps.size == 27
    np.savez_compressed(os.path.join(HERE, "fir_taps.npz"), taps=taps)
    print("fir taps:", taps.size)


def tx_vectors():
    sig = np_oracle.load_reference_simulation("signal")
    s = sig.Signal(f0=16000, f1=19000, fs=44100, T=0.0262, A=20000)          # ChirpGenerator.ipynb cell 1
    H, L, G = s.chirp_orth(), s.chirp_orth(updown="down"), s.silence()
    tone = G
    data = []
    for a in [ord(c) for c in "Hi"]:                                         # ascii(), cell 1
        for i in range(8):
            data.append(H if (a & (0b10000000 >> i)) > 0 else L)
    for t in [H] * 7 + [L] + data + [G] * 12:                                # cell 3
        tone = np.append(tone, t)
    np.savez_compressed(os.path.join(HERE, "tx_vectors.npz"), H=H, L=L, tone_i16=np.real(tone).astype(np.int16),
                        fs=np.array([44100.0]), T=np.array([0.0262]))
    print("tx vectors:", H.size, tone.size)


def wire_excerpt():
    import json
    base = sorted(glob.glob(os.path.join(REF, "agent", "chirp_experiment", "*.raw")))[0][:-4]
    out = {"capture": os.path.basename(base)}
    for ext in ("raw", "flt", "fft"):
        lines = open(base + "." + ext).read().splitlines()
        rows = [1, 2, 3, 700, 1023 if ext == "fft" else 2047]            # 1-based file line = data row + 1 (header is line 0)
        out[ext] = {"header": lines[0], "nlines": len(lines), "rows": {str(r): lines[r] for r in rows}}
    with open(os.path.join(HERE, "wire_excerpt.json"), "w") as f:
        json.dump(out, f, indent=1)
    print("wire_excerpt.json:", out["capture"])


if __name__ == "__main__":
    if "--wire-only" in sys.argv:
        wire_excerpt()
        sys.exit(0)
    if "--tx-only" in sys.argv:
        tx_vectors()
        sys.exit(0)
    tx_vectors()
    wire_excerpt()
    device_triples()
    refsim_vectors()
    fir_taps()
