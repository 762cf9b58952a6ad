"""GPU parity of the two earlier detectors (k_legacy.cu) against the oracle: magnitudes bit-identical,
strengths / levels / codes / decoded bytes exact."""
import numpy as np
import pytest

import synth
import usc
from oracle import pyref as R

pytestmark = pytest.mark.gpu
N = 2048


def _noise_frames(nf, seed, scale=3.0e4):
    return (np.rint(np.random.default_rng(seed).standard_normal((nf, N)) * scale).astype(np.int64) * 256).astype(np.int32)


@pytest.mark.parametrize("nf", [1, 2, 7, 64, 333])
def test_band_magnitudes_bit_exact(nf):
    pcm = _noise_frames(nf, nf)
    pcm[0, :] = 0
    h = usc.Handle()
    d = h.buffer(pcm)
    m = h.empty(4 * nf * 512)
    h.band_magnitudes(d, usc.PCM_I32, nf, m)
    h.sync()
    want = R.legacy_magnitudes(pcm)[:, :512]
    assert np.array_equal(m.to_numpy(np.float32).reshape(nf, 512).view(np.uint32), want.view(np.uint32))
    # float PCM takes the same path
    df = h.buffer(pcm.astype(np.float32))
    h.band_magnitudes(df, usc.PCM_F32, nf, m)
    h.sync()
    assert np.array_equal(m.to_numpy(np.float32).reshape(nf, 512).view(np.uint32), want.view(np.uint32))
    h.close()


def test_onoff_detector_matches_oracle_and_decodes():
    msgs = [b"OK", b"Hi", b"\x00\xff", b"zz"]
    streams = [synth.make_onoff_stream(m, snr_db=snr, seed=30 + i)[0] for i, (m, snr) in enumerate(zip(msgs, (20.0, 20.0, 10.0, 0.0)))]
    F = min(s.shape[0] for s in streams)
    pcm = np.stack([s[:F] for s in streams])                    # [S, F, N]
    S = len(msgs)
    h = usc.Handle()
    cfg = usc.OnOffConfig()
    usc.load().usc_onoff_default_config(__import__("ctypes").byref(cfg))
    cfg.magnitude_threshold = 3000.0 * 256                      # words are x256
    d = h.buffer(pcm)
    st, lv, ch, nc, er = h.empty(2 * S * F), h.empty(S * F), h.empty(S * 8), h.empty(4 * S), h.empty(4 * S)
    h.onoff_detect(d, usc.PCM_I32, S, F, cfg, st, lv, ch, 8, nc, er)
    h.sync()
    mag = R.legacy_magnitudes(pcm.reshape(S * F, N))
    ws, wl, band = R.onoff_detect(mag, mag_threshold=cfg.magnitude_threshold)
    assert band == (446, 499)
    assert np.array_equal(st.to_numpy(np.uint16), ws) and np.array_equal(lv.to_numpy(np.int8), wl)
    chars, ncs, ers = ch.to_numpy(np.uint8).reshape(S, 8), nc.to_numpy(np.uint32), er.to_numpy(np.uint32)
    for s in range(S):
        out, n, e = R.onoff_decode(wl[s * F:(s + 1) * F], cap=8)
        assert bytes(chars[s, :min(n, 8)]) == out and ncs[s] == n and ers[s] == e, s
    assert bytes(chars[0, :2]) == b"OK" and bytes(chars[1, :2]) == b"Hi"
    # decode only (no per-frame outputs requested)
    h.onoff_detect(d, usc.PCM_I32, S, F, cfg, chars=ch, cap=8, nchars=nc)
    h.sync()
    assert bytes(ch.to_numpy(np.uint8).reshape(S, 8)[0, :2]) == b"OK"
    h.close()


@pytest.mark.parametrize("tolerance", [0, 1])
def test_fsk_detector_matches_oracle_and_decodes(tolerance):
    msgs = [b"Hello", b"World", b"\x0f\xf0abc", b"12345"]
    streams = [synth.make_fsk_stream(m, snr_db=snr, seed=40 + i)[0] for i, (m, snr) in enumerate(zip(msgs, (20.0, 20.0, 10.0, 3.0)))]
    F = min(s.shape[0] for s in streams)
    pcm = np.stack([s[:F] for s in streams])
    S = len(msgs)
    h = usc.Handle()
    cfg = usc.FskConfig()
    usc.load().usc_fsk_default_config(__import__("ctypes").byref(cfg))
    cfg.magnitude_threshold = 5000.0 * 256
    cfg.tolerance = tolerance
    d = h.buffer(pcm)
    cd, mg, fr = h.empty(S * F), h.empty(4 * S * F), h.empty(4 * S * F)
    ch, nc, ns, ne = h.empty(S * 16), h.empty(4 * S), h.empty(4 * S), h.empty(4 * S)
    h.fsk_detect(d, usc.PCM_I32, S, F, cfg, cd, mg, fr, ch, 16, nc, ns, ne)
    h.sync()
    mag = R.legacy_magnitudes(pcm.reshape(S * F, N))
    wc, wm, wf = R.fsk_codes(mag, tolerance=tolerance, mag_threshold=cfg.magnitude_threshold)
    assert np.array_equal(cd.to_numpy(np.uint8), wc)
    assert np.array_equal(mg.to_numpy(np.float32).view(np.uint32), wm.view(np.uint32))
    assert np.array_equal(fr.to_numpy(np.float32).view(np.uint32), wf.view(np.uint32))
    chars = ch.to_numpy(np.uint8).reshape(S, 16)
    for s in range(S):
        out, n, sof, eof = R.fsk_parse(wc[s * F:(s + 1) * F], cap=16)
        assert bytes(chars[s, :min(n, 16)]) == out and nc.to_numpy(np.uint32)[s] == n
        assert ns.to_numpy(np.uint32)[s] == sof and ne.to_numpy(np.uint32)[s] == eof
    assert bytes(chars[0, :5]) == b"Hello" and bytes(chars[1, :5]) == b"World"
    h.close()


def test_legacy_argument_checks():
    h = usc.Handle(usc.default_config(n=4096))
    d = h.empty(4 * 4096)
    with pytest.raises(usc.UscError):
        h.band_magnitudes(d, usc.PCM_I32, 1, d)                 # n = 2048 only
    h.close()
    h = usc.Handle(usc.default_config(fs=48000.0))
    d = h.empty(4 * N)
    with pytest.raises(usc.UscError):
        h.onoff_detect(d, usc.PCM_I32, 1, 1)                    # 19 kHz sits above bin 511 at 48 kHz
    h.close()
