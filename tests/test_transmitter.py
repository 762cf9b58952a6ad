"""The transmitter side (include/usc_tx.h, host/usc_tx.c) against the reference's own
output: generator/ChirpGenerator.ipynb cells 1-3 run through simulation/signal.py gave the int16 tone of
"Hi" in tests/golden/tx_vectors.npz.  Then the loop-back of EXPERIMENT4: that 44.1 kHz track, rendered at
the receiver's 78.125 kHz by the resampler's oracle twin, decodes through the oracle receiver."""
import ctypes as C
import os

import numpy as np
import pytest

from oracle import pyref as R

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(os.path.dirname(HERE), "ultrasonic-communication_b200", "libusc_wire.so")
FS, F0, F1, T, A = 44100.0, 16000.0, 19000.0, 0.0262, 20000.0


@pytest.fixture(scope="module")
def tx():
    L = C.CDLL(LIB)
    L.usc_tx_frame_len.restype = C.c_size_t
    L.usc_tx_frame_i16.restype = C.c_long
    L.usc_wav_write_i16.restype = C.c_long
    L.usc_wav_read_i16.restype = C.c_long
    return L


@pytest.fixture(scope="module")
def golden():
    return np.load(os.path.join(HERE, "golden", "tx_vectors.npz"))


def frame(tx, msg, guard=12):
    n = tx.usc_tx_frame_len(C.c_double(FS), C.c_double(T), C.c_uint32(len(msg)), C.c_uint32(guard))
    out = np.zeros(n, np.int16)
    m = (C.c_uint8 * max(len(msg), 1))(*msg)
    got = tx.usc_tx_frame_i16(C.c_double(FS), C.c_double(F0), C.c_double(F1), C.c_double(T), C.c_double(A), m,
                              C.c_uint32(len(msg)), C.c_uint32(guard), out.ctypes.data_as(C.c_void_p), C.c_size_t(n))
    assert got == n
    return out


def test_symbols_and_tone_equal_the_reference_transmitter(tx, golden):
    assert tx.usc_tx_symbol_len(C.c_double(FS), C.c_double(T)) == golden["H"].size == 1155
    for kind, key in ((1, "H"), (2, "L")):
        s = np.zeros(1155)
        assert tx.usc_tx_symbol(C.c_double(FS), C.c_double(F0), C.c_double(F1), C.c_double(T), C.c_double(A), C.c_int(kind),
                                s.ctypes.data_as(C.c_void_p), C.c_uint32(1155)) == 1155
        assert np.abs(s - golden[key]).max() <= 1e-9 * A           # libm vs numpy: last-bit differences at most
    tone = frame(tx, b"Hi")
    d = tone.astype(int) - golden["tone_i16"].astype(int)
    assert tone.size == golden["tone_i16"].size and np.abs(d).max() <= 1 and np.count_nonzero(d) <= tone.size // 1000
    assert tx.usc_tx_symbol(C.c_double(FS), C.c_double(F0), C.c_double(F1), C.c_double(T), C.c_double(A), C.c_int(3), None, C.c_uint32(0)) == -1


def test_wav_round_trip(tx, tmp_path):
    tone = frame(tx, b"ok")
    path = str(tmp_path / "ChirpTone.wav").encode()
    assert tx.usc_wav_write_i16(path, C.c_uint32(44100), tone.ctypes.data_as(C.c_void_p), C.c_size_t(tone.size)) == tone.size
    from scipy.io import wavfile                                   # the reader the reference's tooling would use
    fs, x = wavfile.read(path.decode())
    assert fs == 44100 and x.dtype == np.int16 and np.array_equal(x, tone)
    back = np.zeros(tone.size, np.int16)
    fs2 = C.c_uint32(0)
    assert tx.usc_wav_read_i16(path, C.byref(fs2), back.ctypes.data_as(C.c_void_p), C.c_size_t(back.size)) == tone.size
    assert fs2.value == 44100 and np.array_equal(back, tone)
    wavfile.write(str(tmp_path / "scipy.wav"), 44100, tone)        # and a file written by scipy reads back here
    assert tx.usc_wav_read_i16(str(tmp_path / "scipy.wav").encode(), C.byref(fs2), back.ctypes.data_as(C.c_void_p), C.c_size_t(back.size)) == tone.size
    assert np.array_equal(back, tone)
    assert tx.usc_wav_read_i16(b"/nonexistent.wav", None, None, C.c_size_t(0)) == -2


def test_resampler_twin_properties():
    # unit ratio is the identity (x256); a tone keeps its frequency and amplitude at 3125/1764
    x = (np.random.default_rng(1).integers(-20000, 20000, 5000)).astype(np.int16)
    assert np.array_equal(R.resample_i16_to_pcm(x, 1, 1), x.astype(np.int32) * 256)
    n = 44100
    t = np.arange(n) / 44100.0
    tone = np.rint(15000 * np.sin(2 * np.pi * 17000.0 * t)).astype(np.int16)
    y = R.resample_i16_to_pcm(tone, 3125, 1764) / 256.0
    assert y.size == 78125
    mid = y[2000:-2000]
    tt = (np.arange(y.size) / 78125.0)[2000:-2000]
    ref = 15000 * np.sin(2 * np.pi * 17000.0 * tt)
    assert np.abs(mid - ref).max() < 0.02 * 15000                  # 32-tap Hann-windowed sinc at 0.77 of Nyquist


def test_reference_transmitter_track_decodes_at_the_receiver_rate(golden):
    """EXPERIMENT4 loop-back: the reference's own 44.1 kHz tone for "Hi" -> 78.125 kHz -> receiver -> 'Hi\\n'."""
    lead = np.zeros(44100, np.int16)                               # one second of silence: the noise-floor warm-up
    track = np.concatenate([lead, golden["tone_i16"], np.zeros(8000, np.int16)])
    rng = np.random.default_rng(3)
    track = (track + np.rint(rng.standard_normal(track.size) * 300)).astype(np.int16)   # a noise floor for the SNR estimate
    pcm = R.resample_i16_to_pcm(track, 3125, 1764)
    nframes = pcm.size // 2048
    out, st = R.receiver_run(R.RefReceiver(), pcm[:nframes * 2048].reshape(nframes, 2048), cap=16)
    assert out == b"Hi\n" and st.lock_frame > 20


def test_iq_transmitter_symbol_is_the_notebooks_chirp_iq(tx):
    """generator/ChirpGeneratorIQmodulation.ipynb cells 2-5: chirp_iq() at fs 44100, T 0.0205, BW 3000, carrier 18000,
    amplitude 20000, phase -pi/2 — restated with numpy exactly as the cell is written — and the notebook's own known
    answer (cell 8 prints spectral peaks at 16830.2 and 19171.8 Hz for both symbols)."""
    fs, Tq, BWq, FC, Aq = 44100.0, 0.0205, 3000.0, 18000.0, 20000.0
    n = int(Tq * fs)
    assert tx.usc_tx_symbol_len(C.c_double(fs), C.c_double(Tq)) == n == 904
    t = np.linspace(0, Tq, n)
    k = float(BWq) / float(Tq)
    for kind, fb in ((1, -BWq / 2 + k * t / 2.0), (2, BWq / 2 - k * t / 2.0)):
        want = np.cos((2.0 * np.pi * (FC + fb) * t) + (-np.pi / 2.0)) * Aq
        s = np.zeros(n)
        assert tx.usc_tx_symbol_iq(C.c_double(fs), C.c_double(BWq), C.c_double(FC), C.c_double(Tq), C.c_double(Aq),
                                   C.c_double(-np.pi / 2.0), C.c_int(kind), s.ctypes.data_as(C.c_void_p), C.c_uint32(n)) == n
        assert np.abs(s - want).max() <= 1e-9 * Aq
        a = np.abs(np.fft.fftshift(np.fft.fft(s)))
        freq = np.fft.fftshift(np.fft.fftfreq(n, 1 / fs))
        pos = freq > 0
        # the two spectral horns of a linear chirp: the notebook's peakutils output 16830.2 / 19171.8 Hz
        lo = freq[pos][np.argmax(np.where(freq[pos] < FC, a[pos], 0))]
        hi = freq[pos][np.argmax(np.where(freq[pos] > FC, a[pos], 0))]
        assert abs(lo - 16830.19911504) < 1e-6 and abs(hi - 19171.7920354) < 1e-6
    assert tx.usc_tx_symbol_iq(C.c_double(fs), C.c_double(BWq), C.c_double(FC), C.c_double(Tq), C.c_double(Aq), C.c_double(0.0),
                               C.c_int(0), s.ctypes.data_as(C.c_void_p), C.c_uint32(n)) == n and not s.any()


def test_iq_generator_twin_feeds_the_iq_demodulator(fir_taps):
    """The I/Q transmitter through the counter-based generator (CPU twin of usc_synth_iq_frames) decodes through the
    oracle's I/Q demodulator: in the simulation notebook's sense (carrier - fb) bit 1 = up."""
    taps = fir_taps.astype(np.float32)[::-1].copy()
    q = R.RefIq(taps)
    pcm, bits = R.synth_iq_frames(11, 0, 24, 18000.0, 3000.0, -1, 0.0, 2.0e4, 6000.0)
    assert np.all(pcm % 256 == 0) and 4 < bits.sum() < 20
    mu, iu, md, idn = q.demod(pcm)
    assert np.array_equal((~(md > mu)).astype(np.uint8), bits)
    # the generator notebook's sense (carrier + fb) mirrors the sweep: the same demodulator then reads the complement
    pcm2, bits2 = R.synth_iq_frames(11, 0, 24, 18000.0, 3000.0, +1, -np.pi / 2, 2.0e4, 6000.0)
    mu, iu, md, idn = q.demod(pcm2)
    assert np.array_equal(bits2, bits) and np.array_equal((md > mu).astype(np.uint8), bits2)
