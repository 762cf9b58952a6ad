"""T2: the fp32 C oracle against float64 numpy on seeded synthetic frames, and the semantics the
oracle DEFINES where the reference has hazards or tie rules."""
import numpy as np
import pytest

import synth
from oracle import np_oracle as NP
from oracle import pyref as R

N = 2048
RTOL = 1e-4          # north-star tolerance for peak magnitudes / spectra (fp32)


@pytest.fixture(scope="module")
def rx():
    return R.RefReceiver()


@pytest.fixture(scope="module")
def frames():
    return synth.make_frames(256)


@pytest.mark.parametrize("n", [16, 64, 512, 1024, 2048, 4096])
def test_cfft_vs_numpy(n):
    rng = np.random.default_rng(n)
    x = rng.standard_normal(2 * n).astype(np.float32)
    c = R.Cfft(n)
    y = c(x)
    ref = np.fft.fft(x[0::2].astype(np.float64) + 1j * x[1::2])
    assert np.abs((y[0::2] + 1j * y[1::2]) - ref).max() <= 2e-6 * np.abs(ref).max()
    assert np.abs(c(y, inverse=True) - x).max() <= 4e-6          # forward unscaled, inverse 1/N


@pytest.mark.parametrize("n", [32, 128, 2048, 4096, 8192, 65536])
def test_rfft_packed_layout_vs_numpy(n):
    rng = np.random.default_rng(n + 1)
    x = rng.standard_normal(n).astype(np.float32)
    r = R.Rfft(n)
    y = r(x)
    want = NP.pack_rfft(np.fft.rfft(x.astype(np.float64)), n)
    assert np.abs(y - want).max() <= 2e-6 * np.abs(want).max()
    assert np.abs(r(y, inverse=True) - x).max() <= 4e-6          # irfft(rfft(x)) == x


def test_rfft_init_argument_error():
    with pytest.raises(ValueError):
        R.Rfft(1000)
    with pytest.raises(ValueError):
        R.Rfft(16)


def test_receiver_chain_vs_float64(rx, frames):
    """pipeline() mags of the fp32 oracle within 1e-4 of float64 on the same tables; arg-max bins
    equal except where the two best float64 candidates are within fp32 noise (classified)."""
    pcm, _ = frames
    up, hann = rx.table("up_chirp"), rx.table("hann")
    near_ties = 0
    for f in range(64):
        x = pcm[f].astype(np.float32)
        got = rx.pipeline(x, up=True)
        want = NP.receiver_mags_f64(x, up, hann)
        assert np.all(got[N // 2:] == 0.0)                      # hazard H1 defined: zeros above
        w = want[:156]
        assert np.abs(got[:156] - w).max() <= RTOL * w.max()
        gi, wi = int(np.argmax(got[:156])), int(np.argmax(w))
        if gi != wi:
            assert abs(w[gi] - w[wi]) <= 1e-5 * w[wi]
            near_ties += 1
    assert near_ties <= 1


def test_dsp_history_fields(rx, frames):
    pcm, _ = frames
    fifo = np.zeros(3 * N, np.float32)
    fifo[N:2 * N] = pcm[3]
    h = rx.dsp(fifo, N, mag_mean=1.0e8, up=True)
    mags = rx.pipeline(pcm[3].astype(np.float32), up=True)
    assert h.max_idx_right == int(np.argmax(mags[:156])) and h.mag_max_right == mags[:156].max()
    assert h.mag_max_left == 0.0 and h.max_idx_left == 1892       # H1: left window is zeros
    assert h.mag_max == h.mag_max_right and h.max_idx == h.max_idx_right
    assert h.max_freq == int(78125 * h.max_idx // 2048)           # idx2freq, main.c:154-160
    assert h.max_freq_left == -(78125 * (2048 - 1892) // 2048)
    assert h.snr == np.float32((np.float32(h.mag_max) - np.float32(1.0e8)) / np.float32(1.0e8))


def test_dsp_is_shift_of_aligned(rx, frames):
    """dsp(sync_position=p) on a fifo == aligned demod of fifo[p:p+N] for every offset."""
    pcm, _ = frames
    fifo = np.concatenate([pcm[0], pcm[1], pcm[2]]).astype(np.float32)
    for p in (0, 1, 255, 256, 1024 + 3 * 256, 4096):
        h = rx.dsp(fifo, p, 1.0, up=False)
        mu, iu, md, idn = rx.demod_frames(fifo[p:p + N].reshape(1, N).copy())
        assert (h.mag_max, h.max_idx) == (md[0], idn[0])


def test_demod_i32_equals_f32(rx, frames):
    """(float) buf[i] (main.c:663-665) is the whole PCM scaling: int32 and pre-cast float agree."""
    pcm, _ = frames
    a = rx.demod_frames(pcm)
    b = rx.demod_frames(pcm.astype(np.float32))
    for x, y in zip(a, b):
        assert np.array_equal(x, y)


def test_symbol_decision_quality(rx):
    """At +10 dB every symbol of the synthetic stream is decided correctly (up iff bit 1)."""
    pcm, bits = synth.make_frames(128, snr_db=10.0)
    mu, iu, md, idn = rx.demod_frames(pcm, nthreads=4)
    assert np.array_equal((~(md > mu)).astype(np.uint8), bits)


def test_arm_max_first_occurrence_and_mean_order():
    assert R.arm_max_f32(np.array([1, 5, 5, 2, 5], np.float32)) == (np.float32(5), 1)
    assert R.arm_max_f32(np.array([-3, -3], np.float32)) == (np.float32(-3), 0)
    v = np.array([1e8, 1, -1e8, 1, 1, 1, 1, 1], np.float32)
    s = np.float32(0)
    for x in v:
        s = np.float32(s + x)
    assert R.arm_mean_f32(v) == np.float32(s / np.float32(8))     # sequential sum, one division


def test_cmplx_ops_contracts():
    rng = np.random.default_rng(5)
    a, b = rng.standard_normal(64).astype(np.float32), rng.standard_normal(64).astype(np.float32)
    za, zb = a[0::2] + 1j * a[1::2].astype(np.float64), b[0::2] + 1j * b[1::2].astype(np.float64)
    got = R.arm_cmplx_mult_cmplx_f32(a, b)
    assert np.abs((got[0::2] + 1j * got[1::2]) - za * zb).max() < 1e-6       # no conjugate
    assert np.abs(R.arm_cmplx_mag_f32(a) - np.abs(za)).max() < 1e-6
    r = rng.standard_normal(32).astype(np.float32)
    gr = R.arm_cmplx_mult_real_f32(a, r)
    assert np.array_equal(gr[0::2], a[0::2] * r) and np.array_equal(gr[1::2], a[1::2] * r)


def test_fir_matches_lfilter_and_carries_state(fir_taps):
    """arm_fir_f32 contract (arm_math.h:1194-1214): reversed coefficient order, state carried
    across blocks; vs float64 direct convolution."""
    from scipy.signal import lfilter
    rng = np.random.default_rng(6)
    x = rng.standard_normal(3 * 256).astype(np.float32)
    b = fir_taps.astype(np.float32)
    fir = R.Fir(b[::-1].copy(), 256)
    y = np.concatenate([fir(x[i * 256:(i + 1) * 256]) for i in range(3)])
    want = lfilter(b.astype(np.float64), 1.0, x.astype(np.float64))
    assert np.abs(y - want).max() < 2e-6


def test_compress_chain_vs_float64():
    """compress_chirp (chirp.c:78-83) incl. the packed DC/Nyquist quirk vs float64; the peak lag of a
    delayed up-chirp against H_down is where the float64 chain puts it."""
    c = R.RefCompressor()
    w, H = c.table("window"), c.table("H_down")
    p = R.generate_ref_chirp("T", N, 100000.0, 17000.0, 18000.0, 0.0, float(np.float32(-1.5707963705062866)), 1)
    rng = np.random.default_rng(7)
    for shift in (0, 100, 777):
        x = (np.roll(p, shift) * 20000 + rng.standard_normal(N) * 3000).astype(np.float32)
        got = c.compress(x)
        want = NP.compress_chain_f64(x, w, H, quirk=True)
        assert np.abs(got - want).max() <= RTOL * np.abs(want).max()
        assert int(np.argmax(got)) == int(np.argmax(want))
        noq = NP.compress_chain_f64(x, w, H, quirk=False)
        assert int(np.argmax(noq)) == int(np.argmax(want))      # in-band peak unaffected by the quirk


def test_compress_in_place_aliasing_defined():
    """Hazard H2: the reference calls arm_rfft_fast_f32 with p == pOut; our definition is the
    mathematically correct result either way."""
    r = R.Rfft(256)
    x = np.random.default_rng(8).standard_normal(256).astype(np.float32)
    y = r(x)
    buf = x.copy()
    R.lib().ref_arm_rfft_fast_f32(__import__("ctypes").byref(r.S), buf.ctypes.data_as(R.f32p), buf.ctypes.data_as(R.f32p), 0)
    assert np.array_equal(buf, y)


# ---- state machine (receiver/Src/main.c:417-580) -------------------------------------------------
def test_state_machine_decodes_hello_world(rx):
    """SURVEY §4 validation (3): 40xG lead-in (24-frame mag_stat warm-up), 7xH, L, bits, 12xG at
    +26 dB decodes b'Hello World!' locking at frame 44 with sync_position 2304."""
    pcm = synth.make_stream(b"Hello World!", snr_db=26.0)
    out, st = R.receiver_run(rx, pcm)
    assert out == b"Hello World!\n"
    assert (st.lock_frame, st.lock_position, st.state) == (44, 2304, 0)


def test_state_machine_warmup_blocks_early_detection(rx):
    """mag_stat starts at 1e37 (main.c:321-322): nothing can lock before 12 decisions = 24 frames."""
    pcm = synth.make_stream(b"Hi", snr_db=30.0, lead_in=2)
    out, st = R.receiver_run(rx, pcm)
    assert out == b"" and st.lock_frame == -1


def test_state_machine_offsets_and_noise(rx):
    for off in (0, 256, 700, 1999):
        out, st = R.receiver_run(rx, synth.make_stream(b"OK", snr_db=24.0, start_offset=off))
        assert out == b"OK\n", off
        assert st.lock_position % 256 == 0 and 1024 <= st.lock_position <= 1024 + 7 * 256   # N/2 + k*N/8
    rng = np.random.default_rng(1)
    noise = (np.rint(rng.standard_normal((80, N)) * 5000).astype(np.int64) * 256).astype(np.int32)
    out, st = R.receiver_run(rx, noise)
    assert out == b"" and st.lock_frame == -1


def test_sync_search_grid_matches_dsp(rx):
    """ref_sync_search == dsp(UP) on the 3-frame FIFO at N/2 + turn*N/8 + i*N/4 (main.c:447-451)."""
    pcm = synth.make_stream(b"x", snr_db=10.0, lead_in=3, nframes=12)
    mag, idx = R.sync_search(rx, pcm, 1)
    flat = np.concatenate([np.zeros(2 * N, np.float32), pcm.reshape(-1).astype(np.float32)])
    for t in (0, 1, 5, 8):
        fifo = flat[t * N:(t + 3) * N]
        for i in range(4):
            h = rx.dsp(fifo, N // 2 + (t & 1) * (N // 8) + i * (N // 4), 1.0, up=True)
            assert (mag[t, i], idx[t, i]) == (h.mag_max, h.max_idx)


# ---- I/Q baseband path (config 3) -----------------------------------------------------------------
def test_iq_chain_decides_symbols_and_peaks_at_dc(fir_taps):
    """IQ_modulation.ipynb cells 28-31: the right hypothesis collapses to ~0 Hz, the wrong one to a
    double peak near +-1.6 kHz (outside the +-32-bin window), so up/down is decided by the larger peak."""
    q = R.RefIq(fir_taps.astype(np.float32)[::-1].copy())
    pcm, bits = synth.make_iq_stream(24, snr_db=10.0)
    mu, iu, md, idn = q.demod(pcm)
    assert np.array_equal((~(md > mu)).astype(np.uint8), bits)
    right = np.where(bits == 1, iu, idn)
    assert np.all((right <= 5) | (right >= 1019))
    ratio = np.where(bits == 1, mu / md, md / mu)
    assert ratio.min() > 3.0


def test_iq_fir_state_carries_across_frames(fir_taps):
    """arm_fir_f32 keeps numTaps-1 samples of history between calls (arm_math.h:1194-1214): frame t's
    result depends on the tail of frame t-1."""
    q = R.RefIq(fir_taps.astype(np.float32)[::-1].copy())
    pcm, _ = synth.make_iq_stream(2, snr_db=20.0)
    both = q.demod(pcm)
    alone = q.demod(pcm[1:2])
    assert both[0][1] != alone[0][0]
    assert abs(both[0][1] - alone[0][0]) < 0.05 * both[0][1]
