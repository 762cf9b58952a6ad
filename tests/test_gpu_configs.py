"""The fused kernels under receiver configurations other than the firmware's defaults: sampling rate,
band, sweep time and SNR threshold all enter the tables, the bandwidth2 window and the state machine.
Every combination is checked bit for bit against the oracle built with the same parameters."""
import numpy as np
import pytest

import synth
import usc
from oracle import pyref as R

pytestmark = pytest.mark.gpu
N = 2048

CONFIGS = [
    dict(fs=78125.0, f0=16000.0, f1=19000.0, sweep_T=0.0205),          # the final receiver
    dict(fs=100000.0, f0=16000.0, f1=18000.0, sweep_T=0.02048),        # the earlier 100 kHz set-up (EXPERIMENT3)
    dict(fs=78125.0, f0=17000.0, f1=18000.0, sweep_T=0.0262144),       # narrow band, sweep = frame length
    dict(fs=48000.0, f0=15000.0, f1=19500.0, sweep_T=0.03),            # wide band at a low rate: bandwidth2 = 384
    dict(fs=78125.0, f0=18000.0, f1=18300.0, sweep_T=0.0205),          # very narrow: bandwidth2 < 32
]


@pytest.mark.parametrize("c", CONFIGS, ids=lambda c: "fs%d_%d-%d" % (c["fs"], c["f0"], c["f1"]))
def test_fused_kernels_follow_the_configuration(c):
    h = usc.Handle(usc.default_config(**c))
    rx = R.RefReceiver(**c)
    assert h.geometry() == (rx.rx.bandwidth, rx.bandwidth2, rx.idx_left_zero)
    for t in ("up", "down", "hann"):
        assert np.array_equal(h.table(t), rx.table({"up": "up_chirp", "down": "down_chirp", "hann": "hann"}[t]))
    pcm, _ = synth.make_frames(65, snr_db=0.0, seed_noise=int(c["fs"]) % 97)
    # K1 dual
    want = rx.demod_frames(pcm, nthreads=4)
    got = h.demod_frames_host(pcm)
    for g, w in zip(got[:4], want):
        assert np.array_equal(g.view(np.uint32), w.view(np.uint32))
    # K1 single hypothesis (pair mode), both directions
    d = h.buffer(pcm)
    a, b = h.empty(4 * 65), h.empty(4 * 65)
    h.demod_frames(d, usc.PCM_I32, 65, mag_up=a, idx_up=b)
    h.sync()
    assert np.array_equal(a.to_numpy(np.float32).view(np.uint32), want[0].view(np.uint32)) and np.array_equal(b.to_numpy(np.uint32), want[1])
    h.demod_frames(d, usc.PCM_I32, 65, mag_down=a, idx_down=b)
    h.sync()
    assert np.array_equal(a.to_numpy(np.float32).view(np.uint32), want[2].view(np.uint32)) and np.array_equal(b.to_numpy(np.uint32), want[3])
    if rx.bandwidth2 <= 160:                                            # K4 / K7 serve windows up to 160 bins
        streams = pcm[:64].reshape(2, 32, N)
        ds = h.buffer(streams)
        for K in (1, 3):
            dm, di = h.empty(4 * 2 * 32 * 4), h.empty(4 * 2 * 32 * 4)
            h.sync_search(ds, usc.PCM_I32, 2, 32, 32 * N, K, dm, di)
            h.sync()
            for s in range(2):
                wm, wi = R.sync_search(rx, streams[s], K)
                assert np.array_equal(dm.to_numpy(np.float32).reshape(2, 32, 4)[s].view(np.uint32), wm.view(np.uint32))
                assert np.array_equal(di.to_numpy(np.uint32).reshape(2, 32, 4)[s], wi)
    else:
        with pytest.raises(usc.UscError):
            h.sync_search(h.buffer(pcm[:4]), usc.PCM_I32, 1, 4, 4 * N, 1, a, b)
    h.close()


@pytest.mark.parametrize("thr", [0.5, 2.0, 6.0])
def test_receiver_state_machine_follows_the_threshold(thr):
    """SNR_THRESHOLD (receiver/Inc/main.h:97) decides locking and message ends; 6.0 never locks at this SNR."""
    h = usc.Handle(usc.default_config(snr_threshold=thr))
    rx = R.RefReceiver()
    streams = np.stack([synth.make_stream(b"Hi", snr_db=snr, nframes=80, seed=90 + i, start_offset=300 * i)
                        for i, snr in enumerate((20.0, 8.0, 2.0))])
    out, res = h.receiver_run_host(streams, uart_cap=32)
    for s in range(3):
        want, st = R.receiver_run(rx, streams[s], snr_threshold=thr, cap=32)
        assert out[s] == want and res["lock_frame"][s] == st.lock_frame and res["sync_position"][s] == st.sync_position, s
    h.close()


@pytest.mark.parametrize("fs,f1,f2", [(100000.0, 17000.0, 18000.0), (78125.0, 16000.0, 17500.0), (100000.0, 15000.0, 19000.0)])
@pytest.mark.parametrize("F", [1, 2, 63])
def test_compression_follows_the_configuration(fs, f1, f2, F):
    """K2 under other sampling rates / bands, odd and even batches, int32 and float PCM: peak value, lag and the
    compressed frames bit for bit (experiments/chirp_compression_time_domain chain)."""
    hc = usc.Handle(usc.default_config(fs=fs, f0=f1, f1=f2, chirp_variant=usc.CHIRP_T, window=usc.HANN_SYMMETRIC))
    c = R.RefCompressor(fs=fs, f1=f1, f2=f2)
    assert np.array_equal(hc.table("H_down").view(np.uint32), c.table("H_down").view(np.uint32))
    rng = np.random.default_rng(int(fs + f1) % 1000 + F)
    p = hc.table("up")
    pcm = np.stack([np.rint(np.roll(p, int(rng.integers(0, N))) * 15000 + rng.standard_normal(N) * 9000) for _ in range(F)])
    pcm = (pcm.astype(np.int64) * 256).astype(np.int32)
    d_out, d_v, d_i = hc.empty(4 * F * N), hc.empty(4 * F), hc.empty(4 * F)
    for fmt, data in ((usc.PCM_I32, pcm), (usc.PCM_F32, pcm.astype(np.float32))):
        d = hc.buffer(data)
        hc.compress_chirp(d, fmt, F, False, d_out, d_v, d_i)
        hc.sync()
        wv, wi = c.compress_frames(pcm, use_up=False, nthreads=2)
        assert np.array_equal(d_i.to_numpy(np.uint32), wi)
        assert np.array_equal(d_v.to_numpy(np.float32).view(np.uint32), wv.view(np.uint32))
        want = np.stack([c.compress(pcm[i].astype(np.float32), False) for i in range(F)])
        assert np.array_equal(d_out.to_numpy(np.float32).reshape(F, N).view(np.uint32), want.view(np.uint32))
    hc.close()


def test_iq_path_takes_float_pcm_too(fir_taps):
    taps = fir_taps.astype(np.float32)[::-1].copy()
    h = usc.Handle()
    h.iq_init(18000.0, 3000.0, taps, 32)
    pcm = np.stack([synth.make_iq_stream(9, snr_db=5.0, seed_bits=s, seed_noise=100 + s)[0] for s in range(2)])
    outs = []
    for fmt, data in ((usc.PCM_I32, pcm), (usc.PCM_F32, pcm.astype(np.float32))):
        d = h.buffer(data)
        o = [h.empty(4 * 18) for _ in range(4)]
        b = h.empty(18)
        h.iq_demod(d, fmt, 2, 9, 9 * N, o[0], o[1], o[2], o[3], b)
        h.sync()
        outs.append([x.to_numpy(np.uint32) for x in o] + [b.to_numpy(np.uint8)])
    for a, b2 in zip(*outs):
        assert np.array_equal(a, b2)
    q = R.RefIq(taps)
    want = q.demod(pcm[1])
    assert np.array_equal(outs[0][0][9:], want[0].view(np.uint32)) and np.array_equal(outs[0][1][9:], want[1])
    h.close()
