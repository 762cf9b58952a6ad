"""K6 (GPU): the receiver chain on frame lengths other than 2048 — BASELINE config 5's 8192 / 16384 /
65536-point chirp frames at low SNR (CMSIS tops out at 4096, arm_const_structs.h:49-57, so these
sizes have a numpy/C-oracle reference only) — and the FFT operators at 32768/65536 points."""
import numpy as np
import pytest

import synth
import usc
from oracle import np_oracle as NP
from oracle import pyref as R

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n,frames,snr", [(256, 33, 0.0), (1024, 17, -5.0), (4096, 9, -10.0), (8192, 9, -15.0),
                                          (16384, 5, -15.0), (32768, 3, -15.0), (32768, 71, -20.0), (65536, 3, -20.0), (65536, 77, -15.0),
                                          # eight frames and more per cluster: the split-phase closing barrier of the cluster kernel
                                          # (a cluster's next level 0 overlapping its peers' split) is exercised on every frame but the first
                                          (32768, 530, -20.0), (65536, 270, -15.0)])
def test_long_frame_demod_bit_exact(n, frames, snr):
    h = usc.Handle(usc.default_config(n=n))
    rx = R.RefReceiver(n=n)
    assert h.geometry() == (rx.rx.bandwidth, rx.bandwidth2, rx.idx_left_zero)
    assert np.array_equal(h.table("up"), rx.table("up_chirp")) and np.array_equal(h.table("hann"), rx.table("hann"))
    pcm, bits = synth.make_frames(frames, snr_db=snr, n=n, seed_noise=n)
    want = rx.demod_frames(pcm, nthreads=4)
    got = h.demod_frames_host(pcm)
    for g, w in zip(got[:4], want):
        assert np.array_equal(g.view(np.uint32), w.view(np.uint32))
    assert np.array_equal(got[4], (~(want[2] > want[0])).astype(np.uint8))
    # float64 bound on the same tables (1e-4 on the peak magnitudes)
    up, hann = rx.table("up_chirp"), rx.table("hann")
    for f in range(min(frames, 3)):
        m64 = NP.receiver_mags_f64(pcm[f].astype(np.float32), up, hann)[:rx.bandwidth2]
        assert abs(got[0][f] - m64.max()) <= 1e-4 * m64.max()
    h.close()


@pytest.mark.parametrize("n", [32768, 65536])
def test_rfft_operator_large(n):
    h = usc.Handle()
    rng = np.random.default_rng(n)
    x = (rng.standard_normal((2, n)) * 1e3).astype(np.float32)
    d, o = h.buffer(x), h.empty(x.nbytes)
    h.arm_rfft_fast_f32(n, d, o, 0, 2)
    h.sync()
    got = o.to_numpy(np.float32).reshape(2, n)
    r = R.Rfft(n)
    for i in range(2):
        assert np.array_equal(got[i].view(np.uint32), r(x[i]).view(np.uint32))
    with pytest.raises(usc.UscError):
        h.arm_rfft_fast_f32(n, d, o, 1, 2)                  # inverse only up to 16384
    h.close()


@pytest.mark.parametrize("n", [8192, 16384, 32768])
def test_cfft_operator_large(n):
    h = usc.Handle()
    rng = np.random.default_rng(n + 1)
    x = rng.standard_normal((2, 2 * n)).astype(np.float32)
    d = h.buffer(x)
    h.arm_cfft_f32(n, d, 0, 2)
    h.sync()
    got = d.to_numpy(np.float32).reshape(2, 2 * n)
    c = R.Cfft(n)
    for i in range(2):
        assert np.array_equal(got[i].view(np.uint32), c(x[i]).view(np.uint32))
    h.close()
