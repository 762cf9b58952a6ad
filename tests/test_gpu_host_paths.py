"""Host-buffer entry points (GPU): usc_demod_frames_host on long frames (K6), usc_iq_demod_host (K5) and
usc_receiver_run_host (K7) cut their input into chunks that flow through three stream lanes; whatever the chunk
size, results must equal the device-pointer call on the whole input (which the other GPU tests pin to the oracle)."""
import numpy as np
import pytest

import synth
import usc
from oracle import pyref as R

pytestmark = pytest.mark.gpu
N = 2048


@pytest.mark.parametrize("n,frames,chunk", [(8192, 37, 8), (16384, 11, 8), (65536, 5, 32), (4096, 50, 7)])
def test_long_frames_host_path(n, frames, chunk):
    h = usc.Handle(usc.default_config(n=n))
    rx = R.RefReceiver(n=n)
    pcm, _ = synth.make_frames(frames, snr_db=-10.0, n=n, seed_noise=n + 1)
    want = rx.demod_frames(pcm, nthreads=4)
    h.host_workspace(chunk)                                    # in units of 8 KB: chunk*2048/n frames per chunk (at least one)
    mu, md = np.empty(frames, np.float32), np.empty(frames, np.float32)
    iu, idn = np.empty(frames, np.uint32), np.empty(frames, np.uint32)
    bit = np.empty(frames, np.uint8)
    h.demod_frames_hostbuf(pcm, usc.PCM_I32, frames, mu, iu, md, idn, bit)
    for g, w in zip((mu, iu, md, idn), want):
        assert np.array_equal(g.view(np.uint32), w.view(np.uint32))
    assert np.array_equal(bit, (~(want[2] > want[0])).astype(np.uint8))
    h.close()


def test_host_path_refuses_lengths_without_a_fused_kernel():
    h = usc.Handle(usc.default_config(n=1024))
    pcm = np.zeros((2, 1024), np.int32)
    o = np.empty(2, np.float32)
    with pytest.raises(usc.UscError) as e:
        h.demod_frames_hostbuf(pcm, usc.PCM_I32, 2, o, None, None, None, None)
    assert e.value.code == usc.USC_ERR_ARGUMENT
    h.close()


@pytest.mark.parametrize("chunk,pad", [(12, 0), (5, 64), (4096, 2)])
def test_iq_host_path_equals_device_path(fir_taps, chunk, pad):
    taps = fir_taps.astype(np.float32)[::-1].copy()
    h = usc.Handle()
    h.iq_init(18000.0, 3000.0, taps, 32)
    S, F = 7, 5
    stride = F * N + pad
    pcm = np.zeros((S, stride), np.int32)
    for s in range(S):
        pcm[s, :F * N] = synth.make_iq_stream(F, snr_db=5.0 * s - 10.0, seed_bits=90 + s, seed_noise=95 + s)[0].reshape(-1)
    d = h.buffer(pcm)
    o = [h.empty(4 * S * F) for _ in range(4)]
    b = h.empty(S * F)
    h.iq_demod(d, usc.PCM_I32, S, F, stride, o[0], o[1], o[2], o[3], b)
    h.sync()
    want = [o[0].to_numpy(np.float32), o[1].to_numpy(np.uint32), o[2].to_numpy(np.float32), o[3].to_numpy(np.uint32),
            b.to_numpy(np.uint8)]
    h.host_workspace(chunk)
    got = [np.empty(S * F, np.float32), np.empty(S * F, np.uint32), np.empty(S * F, np.float32), np.empty(S * F, np.uint32),
           np.empty(S * F, np.uint8)]
    h.iq_demod_hostbuf(pcm, usc.PCM_I32, S, F, stride, *got)
    for g, w in zip(got, want):
        assert np.array_equal(g.view(np.uint8), w.view(np.uint8))
    q = R.RefIq(taps)                                           # and the oracle on one stream
    w0 = q.demod(pcm[3, :F * N].reshape(F, N))
    assert np.array_equal(got[0].reshape(S, F)[3].view(np.uint32), w0[0].view(np.uint32))
    h.close()


@pytest.mark.parametrize("chunk_frames,cap", [(6, 32), (7, 32), (64, 5), (160, 32), (1000, 32)])
def test_receiver_host_path_equals_one_device_call(chunk_frames, cap):
    """usc_receiver_run_host: time-sliced chunks of all streams, resumed through the carried state; the bytes, the
    byte counts and the final records equal usc_receiver_run on the whole streams (and the oracle's)."""
    msgs = [b"Hello World!", b"B200", b"\x00\xff\x55\xaa", b"ultrasonic", b"A"]
    cfgs = [(26.0, 0), (26.0, 700), (10.0, 1234), (0.0, 300), (20.0, 2047)]
    F = 160
    pcm = np.stack([synth.make_stream(m, snr_db=snr, start_offset=off, seed=100 + i, nframes=F)
                    for i, (m, (snr, off)) in enumerate(zip(msgs, cfgs))])
    S = pcm.shape[0]
    h = usc.Handle()
    want_uart, want_res = h.receiver_run_host(pcm, uart_cap=cap)          # python helper: one device call
    h.host_workspace(chunk_frames * S)
    uart = np.zeros((S, cap), np.uint8)
    res = np.zeros(S, usc.rx_result_dtype)
    h.receiver_run_hostbuf(pcm, usc.PCM_I32, S, F, F * N, uart, cap, res)
    rx = R.RefReceiver()
    for s in range(S):
        assert bytes(uart[s, :min(int(res["nbytes"][s]), cap)]) == want_uart[s], s
        for k in ("state", "sync_position", "lock_frame", "lock_position", "nbytes", "frames_seen", "turn", "sync_cnt"):
            assert res[k][s] == want_res[k][s], (k, s)
        w, st = R.receiver_run(rx, pcm[s])
        assert want_uart[s] == w[:cap] and res["nbytes"][s] == len(w)
    # a second call on the same handle reuses the staging (state is reset per call)
    uart2 = np.zeros((S, cap), np.uint8)
    res2 = np.zeros(S, usc.rx_result_dtype)
    h.receiver_run_hostbuf(pcm, usc.PCM_I32, S, F, F * N, uart2, cap, res2)
    assert np.array_equal(uart, uart2) and np.array_equal(res, res2)
    h.close()
