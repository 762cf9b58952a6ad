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
