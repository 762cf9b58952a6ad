"""K5 (GPU): the I/Q baseband path (BASELINE config 3) against the oracle's operator-by-operator
restatement of experiments/iq_modulation/Src/iq_modem.c + simulation/IQ_modulation.ipynb; exact."""
import numpy as np
import pytest

import synth
import usc
from oracle import pyref as R

pytestmark = pytest.mark.gpu
N = 2048


def test_iq_demod_matches_oracle(fir_taps):
    taps = fir_taps.astype(np.float32)[::-1].copy()
    h = usc.Handle()
    h.iq_init(18000.0, 3000.0, taps, 32)
    q = R.RefIq(taps)
    S, F = 5, 7
    pcm = np.stack([synth.make_iq_stream(F, snr_db=snr, seed_bits=50 + s, seed_noise=60 + s)[0]
                    for s, snr in enumerate((20.0, 10.0, 0.0, -5.0, 15.0))])
    bits = np.stack([synth.make_iq_stream(F, seed_bits=50 + s)[1] for s in range(S)])
    pcm[4] = 0                                                   # a silent stream
    d = h.buffer(pcm)
    o = [h.empty(4 * S * F) for _ in range(4)]
    b = h.empty(S * F)
    h.iq_demod(d, usc.PCM_I32, S, F, F * N, o[0], o[1], o[2], o[3], b)
    h.sync()
    got = [o[0].to_numpy(np.float32).reshape(S, F), o[1].to_numpy(np.uint32).reshape(S, F),
           o[2].to_numpy(np.float32).reshape(S, F), o[3].to_numpy(np.uint32).reshape(S, F)]
    gbit = b.to_numpy(np.uint8).reshape(S, F)
    for s in range(S):
        want = q.demod(pcm[s])
        for g, w in zip(got, want):
            assert np.array_equal(g[s].view(np.uint32), w.view(np.uint32)), s
        assert np.array_equal(gbit[s], (~(want[2] > want[0])).astype(np.uint8))
    assert np.array_equal(gbit[0], bits[0]) and np.array_equal(gbit[1], bits[1])      # decisions at >= 10 dB
    # peaks of the right hypothesis sit at DC +- a few bins (IQ_modulation.ipynb cells 29-30: -38 / +38 Hz)
    iu = got[1][0][bits[0] == 1]
    assert np.all((iu <= 5) | (iu >= 1024 - 5))
    h.close()


def test_iq_requires_init_and_checks_arguments(fir_taps):
    h = usc.Handle()
    d = h.empty(N * 4)
    with pytest.raises(usc.UscError):
        h.iq_demod(d, usc.PCM_I32, 1, 1, N)
    with pytest.raises(usc.UscError):
        h.iq_init(18000.0, 3000.0, np.zeros(100, np.float32), 32)
    h.close()


@pytest.mark.parametrize("ntaps,window,unfused", [(27, 32, True), (5, 8, False), (16, 32, False), (32, 20, False),
                                                  (1, 32, False), (33, 32, False), (27, 48, False)])
def test_iq_variants_match_oracle(fir_taps, monkeypatch, ntaps, window, unfused):
    """Tap counts either side of the register-blocked FIR's limit (32), odd and even history lengths,
    narrower and wider peak windows, and the operator-chain form (USC_IQ_UNFUSED) — all exact."""
    if unfused:
        monkeypatch.setenv("USC_IQ_UNFUSED", "1")
    rng = np.random.default_rng(ntaps)
    taps = np.resize(fir_taps.astype(np.float32)[::-1], ntaps).copy() if ntaps != 27 else fir_taps.astype(np.float32)[::-1].copy()
    taps *= rng.uniform(0.5, 1.5, ntaps).astype(np.float32)
    h = usc.Handle()
    h.iq_init(18000.0, 3000.0, taps, window)
    q = R.RefIq(taps, window_bins=window)
    S, F = 3, 11
    pcm = np.stack([synth.make_iq_stream(F, snr_db=snr, seed_bits=70 + s, seed_noise=80 + s)[0]
                    for s, snr in enumerate((15.0, 0.0, -8.0))])
    d = h.buffer(pcm)
    o = [h.empty(4 * S * F) for _ in range(4)]
    b = h.empty(S * F)
    h.iq_demod(d, usc.PCM_I32, S, F, F * N, o[0], o[1], o[2], o[3], b)
    h.sync()
    got = [o[0].to_numpy(np.float32).reshape(S, F), o[1].to_numpy(np.uint32).reshape(S, F),
           o[2].to_numpy(np.float32).reshape(S, F), o[3].to_numpy(np.uint32).reshape(S, F)]
    for s in range(S):
        want = q.demod(pcm[s])
        for g, w in zip(got, want):
            assert np.array_equal(g[s].view(np.uint32), w.view(np.uint32)), s
    h.close()
