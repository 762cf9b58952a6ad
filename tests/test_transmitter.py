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
