"""K3 (GPU): dsp() of the complex-FFT `synchronization` variant (experiments/synchronization/Src/
main.c:135-213) — the one where the left (negative-frequency) window is meaningful — against the
oracle's operator-by-operator restatement, bit-exact; plus the operator chain itself on the GPU."""
import numpy as np
import pytest

import synth
import usc
from oracle import pyref as R

pytestmark = pytest.mark.gpu
N = 2048


@pytest.fixture(scope="module")
def hs():
    hnd = usc.Handle(usc.default_config(chirp_variant=usc.CHIRP_S))
    yield hnd
    hnd.close()


@pytest.fixture(scope="module")
def rxs():
    return R.RefSyncReceiver()


def test_tables_match(hs, rxs):
    assert np.array_equal(hs.table("up"), rxs.table("up_chirp"))
    assert np.array_equal(hs.table("down"), rxs.table("down_chirp"))
    assert np.array_equal(hs.table("hann"), rxs.table("hann"))
    assert hs.geometry() == (78, 156, 1892)


def test_dsp_complex_variant_bit_exact(hs, rxs):
    S = 37
    pcm, bits = synth.make_frames(3 * S, snr_db=3.0, seed_noise=41, dtype=np.float32)
    fifo = pcm.reshape(S, 3 * N).copy()
    fifo[5] = 0.0                                              # a silent stream
    pos = (np.arange(S, dtype=np.uint32) * 331) % 4097         # arbitrary offsets in [0, 4096]
    mean = np.linspace(1e7, 4e8, S).astype(np.float32)
    d_f, d_p, d_m, d_h = hs.buffer(fifo), hs.buffer(pos), hs.buffer(mean), hs.empty(48 * S)
    left_wins = 0
    for updown in (usc.UP, usc.DOWN):
        hs.dsp(d_f, 3 * N, d_p, d_m, updown, d_h, S)
        hs.sync()
        got = d_h.to_numpy(usc.history_dtype)
        for s in range(S):
            w = rxs.dsp(fifo[s], int(pos[s]), float(mean[s]), up=bool(updown))
            g = got[s]
            assert (g["max_idx"], g["max_idx_left"], g["max_idx_right"]) == (w.max_idx, w.max_idx_left, w.max_idx_right), s
            assert (g["max_freq"], g["max_freq_left"], g["max_freq_right"]) == (w.max_freq, w.max_freq_left, w.max_freq_right)
            for k in ("mag_max", "mag_max_left", "mag_max_right", "mag_mean", "snr"):
                a, b = np.float32(g[k]), np.float32(getattr(w, k))
                assert a.view(np.uint32) == b.view(np.uint32) or (np.isnan(a) and np.isnan(b)), (s, k)
            left_wins += int(g["max_idx"] >= 1892)
    assert left_wins > 5                                       # the left window really is exercised


def test_operator_chain_on_gpu_equals_fused_kernel(hs, rxs):
    """The same chain built from the batched CMSIS-shaped operators (cmplx_mult_cmplx, cmplx_mult_real,
    cfft 2048, cmplx_mag, max) gives the fused kernel's numbers."""
    B = 6
    pcm, _ = synth.make_frames(B, snr_db=5.0, seed_noise=43, dtype=np.float32)
    z = np.zeros((B, 2 * N), np.float32)
    z[:, 0::2] = pcm
    d_z, d_c, d_w = hs.buffer(z), hs.buffer(hs.table("up")), hs.buffer(hs.table("hann"))
    hs.arm_cmplx_mult_cmplx_f32(d_z, 2 * N, d_c, 0, d_z, 2 * N, N, B)
    hs.arm_cmplx_mult_real_f32(d_z, 2 * N, d_w, 0, d_z, 2 * N, N, B)
    hs.arm_cfft_f32(N, d_z, 0, B)
    d_mag = hs.empty(4 * B * N)
    hs.arm_cmplx_mag_f32(d_z, 2 * N, d_mag, N, N, B)
    hs.sync()
    mags = d_mag.to_numpy(np.float32).reshape(B, N)
    d_v, d_i = hs.empty(4 * B), hs.empty(4 * B)
    hs.arm_max_f32(d_mag, N, 156, d_v, d_i, B)
    hs.sync()
    for b in range(B):
        want = rxs.pipeline(z[b], up=True)
        assert np.array_equal(mags[b].view(np.uint32), want.view(np.uint32))
        fifo = np.zeros(3 * N, np.float32)
        fifo[:N] = pcm[b]
        h = rxs.dsp(fifo, 0, 1.0, True)
        assert (d_v.to_numpy(np.float32)[b], d_i.to_numpy(np.uint32)[b]) == (h.mag_max_right, h.max_idx_right)
