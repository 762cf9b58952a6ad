"""K8 (GPU): the overlap-save synchroniser through the C-ABI against its oracle twin (operator chain rfft_4096 .
arm_cmplx_mult_cmplx_f32 . irfft_4096 on 2n-sample windows): bit-identical floats and lags; every stream length and
segment split; float64 bound; exact lag of a delayed template."""
import numpy as np
import pytest

import usc
from oracle import pyref as R

pytestmark = pytest.mark.gpu
N = 2048


@pytest.fixture(scope="module")
def hc():
    h = usc.Handle(usc.default_config(fs=100000.0, f0=17000.0, f1=18000.0, chirp_variant=usc.CHIRP_T, window=usc.HANN_SYMMETRIC))
    yield h
    h.close()


def templates(h):
    return R.arm_mult_f32(h.table("up"), h.table("hann")), R.arm_mult_f32(h.table("down"), h.table("hann"))


def run(h, pcm, use_up, want_out=True):
    S, F = pcm.shape[0], pcm.shape[1] // N
    d = h.buffer(pcm)
    nb = S * (F - 1)
    out = h.empty(4 * nb * N) if want_out else None
    mv, mi = h.empty(4 * nb), h.empty(4 * nb)
    h.correlate_os(d, usc.PCM_I32 if pcm.dtype == np.int32 else usc.PCM_F32, S, F, F * N, use_up, out, mv, mi)
    h.sync()
    o = out.to_numpy(np.float32).reshape(S, F - 1, N) if want_out else None
    return o, mv.to_numpy(np.float32).reshape(S, F - 1), mi.to_numpy(np.uint32).reshape(S, F - 1)


@pytest.mark.parametrize("S,F", [(1, 2), (3, 3), (5, 9), (2, 40), (700, 4)])
@pytest.mark.parametrize("dtype", [np.int32, np.float32])
def test_bit_exact_against_the_oracle(hc, S, F, dtype):
    rng = np.random.default_rng(S * 100 + F)
    gu, gd = templates(hc)
    if dtype == np.int32:
        pcm = (rng.integers(-2 ** 15, 2 ** 15, size=(S, F * N)) * 256).astype(np.int32)
    else:
        pcm = (rng.standard_normal((S, F * N)) * 3e4).astype(np.float32)
    for use_up, g in ((False, gd), (True, gu)):
        out, mv, mi = run(hc, pcm, use_up)
        for s in (range(S) if S <= 5 else (0, S // 2, S - 1)):
            wo, wv, wi = R.correlate_os(g, pcm[s].reshape(F, N))
            assert np.array_equal(out[s].view(np.uint32), wo.view(np.uint32)), (s, use_up)
            assert np.array_equal(mi[s], wi) and np.array_equal(mv[s].view(np.uint32), wv.view(np.uint32))
    _, mv2, mi2 = run(hc, pcm, False, want_out=False)                     # peaks only
    assert np.array_equal(mi2, run(hc, pcm, False)[2])


def test_linear_filter_property_and_exact_lag(hc):
    gu, gd = templates(hc)
    F = 6
    rng = np.random.default_rng(3)
    x = (rng.standard_normal((1, F * N)) * 2e4).astype(np.float32)
    out, mv, mi = run(hc, x, False)
    want = np.convolve(x[0].astype(np.float64), gd.astype(np.float64))[N:F * N]
    assert np.abs(out.reshape(-1) - want).max() / np.abs(want).max() < 1e-4
    for d in (1, 777, 2048, 3000, 4 * N + 5):
        y = rng.standard_normal((1, F * N)).astype(np.float32) * 50.0
        y[0, d:d + N] += gd[::-1] * np.float32(1e4)
        _, mv, mi = run(hc, y, False, want_out=False)
        t = d + N - 1
        b, lag = t // N - 1, t % N
        if b < F - 1:
            assert int(np.argmax(mv[0])) == b and int(mi[0, b]) == lag


def test_receiver_handle_and_argument_checks(hc):
    h = usc.Handle()                                                      # variant R tables work as the template too
    g = R.arm_mult_f32(h.table("down"), h.table("hann"))
    pcm, _ = R.synth_frames(3, 0, 8, 2.0e4, 1.0e4)
    out, mv, mi = run(h, pcm.reshape(1, -1), False)
    wo, wv, wi = R.correlate_os(g, pcm)
    assert np.array_equal(out[0].view(np.uint32), wo.view(np.uint32)) and np.array_equal(mi[0], wi)
    d = h.buffer(pcm)
    h.correlate_os(d, usc.PCM_I32, 1, 1, N, False, None, None, None)      # fewer than two frames: nothing to do
    with pytest.raises(usc.UscError):
        h.correlate_os(d.ptr + 4, usc.PCM_I32, 1, 8, 8 * N, False, None, None, None)
    with pytest.raises(usc.UscError):
        h.correlate_os(d, usc.PCM_I32, 1, 8, 4 * N, False, None, None, None)   # stride shorter than the stream
    h.close()
