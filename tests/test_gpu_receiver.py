"""T4 (GPU): the receiver's whole main loop (K7) and the sliding-correlation search grid (K4)
against the oracle's restatement of receiver/Src/main.c:417-580 — decoded UART bytes, lock frame
and sync offsets are integers and must be exact."""
import numpy as np
import pytest

import synth
import usc
from oracle import pyref as R

pytestmark = pytest.mark.gpu
N = 2048


@pytest.fixture(scope="module")
def h():
    hnd = usc.Handle()
    yield hnd
    hnd.close()


@pytest.fixture(scope="module")
def rx():
    return R.RefReceiver()


def _streams():
    msgs = [b"Hello World!", b"B200", b"\x00\xff\x55\xaa", b"ultrasonic", b"A", b""]
    cfgs = [(26.0, 0), (26.0, 700), (10.0, 1234), (0.0, 300), (20.0, 2047), (26.0, 1024), (-5.0, 17), (15.0, 1800)]
    out = []
    for i, (snr, off) in enumerate(cfgs):
        out.append(synth.make_stream(msgs[i % len(msgs)], snr_db=snr, start_offset=off, seed=100 + i, nframes=160))
    return np.stack(out)


def test_hello_world_loopback(h, rx):
    """generator/ChirpGenerator.ipynb frame -> receiver: 'Hello World!' (experiments/EXPERIMENT4.md:29)."""
    pcm = synth.make_stream(b"Hello World!", snr_db=26.0, nframes=160)[None]
    uart, res = h.receiver_run_host(pcm)
    assert uart[0] == b"Hello World!\n"
    assert res["lock_frame"][0] == 44 and res["lock_position"][0] == 2304      # SURVEY §4 validation (3)
    assert res["state"][0] == 0


def test_receiver_run_matches_oracle(h, rx):
    pcm = _streams()
    uart, res = h.receiver_run_host(pcm)
    decoded = 0
    for s in range(pcm.shape[0]):
        want, st = R.receiver_run(rx, pcm[s])
        assert uart[s] == want, s
        assert (res["state"][s], res["sync_position"][s], res["lock_frame"][s], res["lock_position"][s]) == \
               (st.state, st.sync_position, st.lock_frame, st.lock_position), s
        assert (res["turn"][s], res["sync_cnt"][s], res["frames_seen"][s]) == (st.turn, st.sync_cnt, st.frames_seen)
        decoded += len(want)
    assert decoded > 20


def test_receiver_noise_only_never_locks(h, rx):
    rng = np.random.default_rng(3)
    pcm = (np.rint(rng.standard_normal((4, 100, N)) * 3000).astype(np.int64) * 256).astype(np.int32)
    uart, res = h.receiver_run_host(pcm)
    for s in range(4):
        want, st = R.receiver_run(rx, pcm[s])
        assert uart[s] == want == b""
        assert res["lock_frame"][s] == st.lock_frame == -1


def test_receiver_many_streams_and_uart_cap(h, rx):
    """130 streams (more than resident warps per CTA wave, ragged) with a tiny UART capacity: the
    byte count keeps counting, the stored prefix is exact."""
    base = synth.make_stream(b"Hello World!", snr_db=20.0, nframes=160)
    pcm = np.stack([np.roll(base.reshape(-1), 64 * s).reshape(160, N) for s in range(130)])
    pcm[:, 0, :] = 0                                               # wrapped tail -> silence
    uart, res = h.receiver_run_host(pcm, uart_cap=5)
    for s in (0, 1, 37, 129):
        want, st = R.receiver_run(rx, pcm[s])
        assert res["nbytes"][s] == len(want)
        assert uart[s] == want[:5]


@pytest.mark.parametrize("sync_add", [1, 2, 4])
def test_sync_search_matches_oracle(h, rx, sync_add):
    pcm = _streams()[:3, :60]
    S, F = pcm.shape[:2]
    d = h.buffer(pcm)
    d_m, d_i = h.empty(4 * S * F * 4), h.empty(4 * S * F * 4)
    h.sync_search(d, usc.PCM_I32, S, F, F * N, sync_add, d_m, d_i)
    h.sync()
    mag = d_m.to_numpy(np.float32).reshape(S, F, 4)
    idx = d_i.to_numpy(np.uint32).reshape(S, F, 4)
    for s in range(S):
        wm, wi = R.sync_search(rx, pcm[s], sync_add)
        assert np.array_equal(idx[s], wi)
        assert np.array_equal(mag[s].view(np.uint32), wm.view(np.uint32))


def test_sync_search_unaligned_streams_take_the_gathered_form(h, rx):
    """K4 stages a frame's window union by TMA bulk copies when the streams are 16-byte aligned; a PCM pointer or a stream
    stride that is not must give the same bits through the gathered form."""
    pcm = _streams()[:3, :40]
    S, F = pcm.shape[:2]
    stride = F * N + 2                                        # the API asks for 8-byte alignment only: stride = 2 (mod 4) samples
    flat = np.zeros(2 + S * stride, dtype=np.int32)
    for s in range(S):
        flat[2 + s * stride: 2 + s * stride + F * N] = pcm[s].reshape(-1)
    d = h.buffer(flat)
    d_m, d_i = h.empty(4 * S * F * 4), h.empty(4 * S * F * 4)
    h.sync_search(d.ptr + 8, usc.PCM_I32, S, F, stride, 1, d_m, d_i)      # base pointer 8 bytes past a 16-byte boundary
    h.sync()
    mag = d_m.to_numpy(np.float32).reshape(S, F, 4)
    idx = d_i.to_numpy(np.uint32).reshape(S, F, 4)
    for s in range(S):
        wm, wi = R.sync_search(rx, pcm[s], 1)
        assert np.array_equal(idx[s], wi)
        assert np.array_equal(mag[s].view(np.uint32), wm.view(np.uint32))


def test_synchronous_addition_raises_the_peak(h):
    """SynchronousAddition.ipynb cells 5-6: adding K frame-aligned preamble frames grows the
    de-chirped peak ~K-fold while noise grows ~sqrt(K).  Needs a transmitter whose symbol period is
    exactly one frame (N/fs), otherwise successive frames are not phase-coherent."""
    pcm = synth.make_stream(b"", snr_db=-3.0, lead_in=4, guard=30, nframes=40, seed=5, tx_T=N / 78125.0)[None]
    d = h.buffer(pcm)
    peaks = []
    for K in (1, 2, 4):
        d_m, d_i = h.empty(4 * 40 * 4), h.empty(4 * 40 * 4)
        h.sync_search(d, usc.PCM_I32, 1, 40, 40 * N, K, d_m, d_i)
        h.sync()
        peaks.append(d_m.to_numpy(np.float32).reshape(40, 4)[8:11].max())
    assert peaks[1] > 1.6 * peaks[0] and peaks[2] > 2.8 * peaks[0]


def test_receiver_argument_errors(h):
    d = h.empty(2 * N * 4)
    r = h.empty(64)
    with pytest.raises(usc.UscError) as e:
        h.receiver_run(d, usc.PCM_I32, 1, 2, N, None, 0, r)         # stride shorter than the stream
    assert e.value.code == usc.USC_ERR_ARGUMENT
    with pytest.raises(usc.UscError):
        h.receiver_run(d, usc.PCM_I32, 1, 2, 2 * N, None, 0, None)   # nothing to write
    with pytest.raises(usc.UscError):
        h.sync_search(d, usc.PCM_I32, 1, 2, 2 * N, 0, r, r)          # sync_add >= 1


@pytest.mark.parametrize("splits", [(70,), (1, 69), (2, 3, 65), (35, 35), (10, 20, 5, 35), tuple([7] * 10)])
def test_receiver_run_in_chunks_equals_one_call(splits):
    """usc_receiver_run_chunk: any split of the stream into chunks (each carrying two frames of history)
    gives the bytes, lock frame and final state of one call over the whole stream — and of the oracle."""
    F = 70
    msgs = [b"Hi", b"ok", b"zz"]
    streams = np.stack([synth.make_stream(m, snr_db=snr, start_offset=off, nframes=F, seed=70 + i)
                        for i, (m, snr, off) in enumerate(zip(msgs, (20.0, 12.0, 3.0), (0, 777, 1999)))])
    S = len(msgs)
    h = usc.Handle()
    whole, res = h.receiver_run_host(streams, uart_cap=32)
    d_state = h.empty(160 * S)
    h_memset = usc.load().usc_memset
    assert h_memset(h._h, __import__("ctypes").c_void_p(d_state.ptr), 0, __import__("ctypes").c_size_t(160 * S)) == 0
    out = [b""] * S
    t0 = 0
    last = None
    for n in splits:
        carry = min(2, t0)
        chunk = np.ascontiguousarray(streams[:, t0 - carry:t0 + n])          # [S, carry + n, N]
        d = h.buffer(chunk)
        d_u, d_r = h.empty(S * 32), h.empty(S * usc.rx_result_dtype.itemsize)
        h.receiver_run_chunk(d, usc.PCM_I32, S, n, (carry + n) * N, carry, d_state, d_u, 32, d_r)
        h.sync()
        r = d_r.to_numpy(usc.rx_result_dtype)
        u = d_u.to_numpy(np.uint8).reshape(S, 32)
        for s in range(S):
            out[s] += bytes(u[s, :min(int(r["nbytes"][s]), 32)])
        t0 += n
        last = r
    assert t0 == F
    rx = R.RefReceiver()
    for s in range(S):
        assert out[s] == whole[s], (s, out[s], whole[s])
        want, st = R.receiver_run(rx, streams[s], cap=32)
        assert out[s] == want[:32]
        for k in ("state", "sync_position", "lock_frame", "lock_position", "frames_seen", "turn", "sync_cnt"):
            assert last[k][s] == res[k][s], (k, s)
    h.close()
