"""The oracle's restatement of the reference's two earlier detectors:
experiments/chirp (on/off chirp band count + decode()) and experiments/ultracom (18-tone FSK + parser())."""
import numpy as np

import synth
from oracle import pyref as R


def test_front_half_is_the_analyser_chain(device_triples):
    """Same chain as experiments/basic fft(): the magnitudes match the captured .fft files (the device
    captures pin mult∘Hann, rfft, mag, scale to 1e-4)."""
    idx = [i for i, ok in enumerate(device_triples["consistent"]) if ok][:6]
    mag = R.legacy_magnitudes(device_triples["raw"][idx].astype(np.int32))
    for r, i in enumerate(idx):
        dev = device_triples["fft_mag"][i]
        sel = device_triples["fft_freq"][i] >= 1000.0
        big = sel & (dev >= 0.01 * dev[sel].max())
        assert (np.abs(mag[r][big] - dev[big]) / dev[big]).max() < 1e-4


def test_onoff_band_and_thresholds():
    _, _, (lo, hi) = R.onoff_detect(np.zeros((1, 1024), np.float32))
    assert (lo, hi) == (446, 499)                     # first bins at/above 17 kHz and 19 kHz at fs = 78125
    _, _, (lo, hi) = R.onoff_detect(np.zeros((1, 1024), np.float32), fs=100000.0)
    assert (lo, hi) == (349, 390)
    m = np.zeros((3, 1024), np.float32)
    m[0, 446:452] = 3001.0                            # 6 of 54 bins: >= int(54*0.1) = 5 -> HIGH
    m[1, 446:449] = 3001.0                            # 3: between int(54*0.05) = 2 and 5 -> UNKNOWN
    m[2, 446:448] = 3001.0                            # 2 -> LOW
    m[2, 300] = 1e9                                   # outside the band
    s, lv, _ = R.onoff_detect(m)
    assert list(s) == [6, 3, 2] and list(lv) == [1, 0, -1]
    m[0, 446:452] = 3000.0                            # strict '>'
    assert R.onoff_detect(m)[0][0] == 0


def test_onoff_decode_state_machine():
    H, L, U = 1, -1, 0
    frame = lambda byte: [H, H, H] + sum(([H, H] if (byte >> b) & 1 else [L, L] for b in range(7, -1, -1)), []) + [L]
    lv = [L] * 5 + frame(0x48) + [L] * 3 + frame(0x69) + [L] * 4
    out, n, errs = R.onoff_decode(lv)
    assert out == b"Hi" and n == 2 and errs == 0
    # one HIGH then silence: the sync check fires when count reaches FRAME_START with fewer than 2 HIGHs
    out, n, errs = R.onoff_decode([H, L, L, L] + frame(0x41) + [L])
    assert errs == 1 and out == b"A"
    # UNKNOWN at a sampling point leaves the bit at 0 but advances
    lv = frame(0xFF)
    lv[4] = U                                         # first data sampling point (count == 4)
    assert R.onoff_decode(lv + [L])[0] == b"\x7f"


def test_onoff_end_to_end_signal():
    pcm, levels = synth.make_onoff_stream(b"OK")
    mag = R.legacy_magnitudes(pcm)
    s, lv, _ = R.onoff_detect(mag, mag_threshold=3000.0 * 256)      # captured words are x256
    assert np.array_equal(lv > 0, levels > 0)
    assert R.onoff_decode(lv)[0] == b"OK"


def test_fsk_codes_and_parser():
    m = np.zeros((6, 1024), np.float32)
    m[0, 340] = 5001.0; m[1, 344] = 5001.0; m[2, 348] = 6000.0; m[3, 408] = 7000.0
    m[4, 340] = 5000.0                                # strict '>'
    m[5, 352] = 9000.0; m[5, 344] = 5001.0            # end-of-frame outranks a hex digit
    code, mg, fr = R.fsk_codes(m)
    assert list(code) == [0xF0, 0xF1, 0, 15, 0xFF, 0xF1]
    assert mg[2] == 6000.0 and fr[2] == np.float32(349) * np.float32(78125.0) / np.float32(2048)   # frequency[j + 1]
    seq = [0xF0] * 3 + [0xFF] + [4] * 3 + [0xFF] + [8] * 3 + [0xFF] + [6] * 3 + [0xFF] + [9] * 5 + [0xFF] + [0xF1] * 3
    out, n, sof, eof = R.fsk_parse(np.array(seq, np.uint8))
    assert out == b"Hi" and (sof, eof) == (1, 1)
    # two sightings are not enough at TQ_N = 2; digits before a start marker are ignored
    assert R.fsk_parse(np.array([4, 4, 0xFF, 8, 8, 8, 8], np.uint8))[1] == 0


def test_fsk_end_to_end_signal():
    pcm, _ = synth.make_fsk_stream(b"Hello")
    mag = R.legacy_magnitudes(pcm)
    code, _, _ = R.fsk_codes(mag, mag_threshold=5000.0 * 256)
    out, n, sof, eof = R.fsk_parse(code)
    assert out == b"Hello" and sof == 1 and eof == 1
