"""Oracle twin of the overlap-save synchroniser (ref_correlate_os) against float64 numpy: the n valid lags of every
block are the LINEAR filter output of the stream (no wrap-around), within the north-star's 1e-4; a delayed copy of the
template peaks exactly at its delay."""
import numpy as np

from oracle import pyref as R

N = 2048


def template():
    c = R.RefCompressor()
    down = R.generate_ref_chirp("T", N, 100000.0, 17000.0, 18000.0, 0.0, np.float32(-3.14159265358979 / 2.0), False)
    return R.arm_mult_f32(down, c.table("window"))


def test_valid_lags_are_the_linear_filter_output():
    rng = np.random.default_rng(5)
    g = template()
    F = 6
    x = (rng.standard_normal(F * N) * 3e4).astype(np.float32)
    out, mv, mi = R.correlate_os(g, x.reshape(F, N))
    want = np.convolve(x.astype(np.float64), g.astype(np.float64))[N:F * N]            # y[t] = sum_m g[m] x[t - m], t >= n
    got = out.reshape(-1).astype(np.float64)
    assert np.abs(got - want).max() / np.abs(want).max() < 1e-4
    assert np.abs(got - want).max() / np.abs(want).max() < 5e-6                         # what it is
    for b in range(F - 1):
        assert mi[b] == int(np.argmax(out[b])) and mv[b] == out[b].max()


def test_delayed_template_peaks_at_its_delay():
    """x = time-reversed template placed at sample d: the matched filter output peaks where the copy ends, t = d + n - 1"""
    g = template()
    F = 5
    for d in (0, 1, 777, 2047, 2048, 3000, 3 * N - 1):
        x = np.zeros(F * N, np.float32)
        x[d:d + N] = g[::-1] * np.float32(1e4)
        out, mv, mi = R.correlate_os(g, x.reshape(F, N))
        t = d + N - 1
        if t < N:
            continue
        b, lag = t // N - 1, t % N
        assert mi[b] == lag and np.argmax(mv) == b, (d, b, lag, mi, mv)


def test_int32_pcm_is_cast_first():
    rng = np.random.default_rng(6)
    g = template()
    x = (rng.integers(-2 ** 23, 2 ** 23, size=(4, N)) * 256).astype(np.int32)
    a = R.correlate_os(g, x)
    b = R.correlate_os(g, x.astype(np.float32))
    assert np.array_equal(a[0].view(np.uint32), b[0].view(np.uint32)) and np.array_equal(a[2], b[2])
