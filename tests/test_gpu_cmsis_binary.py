"""GPU kernels (through the C-ABI) against the reference's OWN CMSIS-DSP machine code: the expected values in
tests/golden/cmsis_binary_vectors.npz were computed by interpreting the reference's vendored ARM archive
(tools/cmsis_emu/).  Peak bins identical; magnitudes, compressed frames and spectra within the north-star's 1e-4."""
import os

import numpy as np
import pytest

import usc
from oracle import pyref as R            # only for regenerating the seeded input frames (integer generator)

pytestmark = pytest.mark.gpu
TOL = 1e-4
GOLD = os.path.join(os.path.dirname(__file__), "golden", "cmsis_binary_vectors.npz")


@pytest.fixture(scope="module")
def g():
    return np.load(GOLD)


def rel_to_peak(got, want):
    want = np.asarray(want, np.float64)
    return float(np.abs(np.asarray(got, np.float64) - want).max() / np.abs(want).max())


def test_fused_demodulator_against_the_binary(g):
    h = usc.Handle()
    for name, key in (("up", "rx_up_chirp"), ("down", "rx_down_chirp"), ("hann", "rx_hann")):
        assert np.array_equal(h.table(name).view(np.uint32), g[key].view(np.uint32))      # the library's host tables
    pcm = np.concatenate([R.synth_frames(int(s), int(f), int(n), a, sg)[0] for s, f, n, a, sg in g["rx_cases"]])
    assert int(np.bitwise_xor.reduce(pcm.view(np.uint32).ravel())) == int(g["rx_pcm_crc"][0])
    mu, iu, md, idn, bit = h.demod_frames_host(pcm)                                        # K1, both hypotheses
    assert np.array_equal(iu, g["rx_peak_idx"][:, 0]) and np.array_equal(idn, g["rx_peak_idx"][:, 1])
    assert np.abs(mu / g["rx_peak"][:, 0] - 1).max() < TOL and np.abs(md / g["rx_peak"][:, 1] - 1).max() < TOL
    assert np.array_equal(bit, (~(g["rx_peak"][:, 1] > g["rx_peak"][:, 0])).astype(np.uint8))
    # usc_pipeline: whole magnitude spectra
    x = pcm.astype(np.float32)
    for hyp, up in ((0, 1), (1, 0)):
        d, o = h.buffer(x), h.empty(x.nbytes)
        h.pipeline(d, o, up, len(x))
        m = o.to_numpy(np.float32).reshape(len(x), 2048)[:, :1024]
        for f in range(len(x)):
            assert rel_to_peak(m[f, 1:], g["rx_mag"][f, hyp][1:]) < TOL
    h.close()


def test_compression_kernel_against_the_binary(g):
    s, f, n, a, sg = g["cc_case"]
    pcm, _ = R.synth_frames(int(s), int(f), int(n), a, sg, n=2048, fs=100000.0, f0=17000.0, f1=18000.0)
    h = usc.Handle(usc.default_config(fs=100000.0, f0=17000.0, f1=18000.0, chirp_variant=usc.CHIRP_T, window=usc.HANN_SYMMETRIC))
    d = h.buffer(pcm)
    out, mv, mi = h.empty(4 * pcm.size), h.empty(4 * len(pcm)), h.empty(4 * len(pcm))
    h.compress_chirp(d, usc.PCM_I32, len(pcm), False, out, mv, mi)
    h.sync()
    got = out.to_numpy(np.float32).reshape(len(pcm), 2048)
    for k in range(len(pcm)):
        assert rel_to_peak(got[k], g["cc_out"][k]) < TOL
    assert np.array_equal(mi.to_numpy(np.uint32), g["cc_idx"])
    assert np.abs(mv.to_numpy(np.float32) / g["cc_max"] - 1).max() < TOL
    h.close()


def test_transform_operators_against_the_binary(g):
    h = usc.Handle()
    for n in (256, 1024, 4096):
        d, o = h.buffer(g["rfft%d_in" % n]), h.empty(4 * n)
        h.arm_rfft_fast_f32(n, d, o, 0, 1)
        assert rel_to_peak(o.to_numpy(np.float32), g["rfft%d_out" % n]) < TOL
        d = h.buffer(g["rfft%d_out" % n])
        h.arm_rfft_fast_f32(n, d, o, 1, 1)
        assert rel_to_peak(o.to_numpy(np.float32), g["rifft%d_out" % n]) < TOL
    for n in (1024, 2048):
        for inv, key in ((False, "cfft%d_out"), (True, "cifft%d_out")):
            d = h.buffer(g["cfft%d_in" % n])
            h.arm_cfft_f32(n, d, inv, 1)
            assert rel_to_peak(d.to_numpy(np.float32), g[key % n]) < TOL
    a, b = g["op_a"], g["op_b"]
    da, db, do = h.buffer(a), h.buffer(b), h.empty(a.nbytes)
    h.arm_mult_f32(da, 512, db, 512, do, 512, 512, 1)
    assert np.array_equal(do.to_numpy(np.float32).view(np.uint32), g["op_mult"].view(np.uint32))
    h.arm_cmplx_mult_cmplx_f32(da, 512, db, 512, do, 512, 256, 1)
    assert rel_to_peak(do.to_numpy(np.float32), g["op_cmul"]) < 2e-7
    dm = h.empty(4 * 256)
    h.arm_cmplx_mag_f32(da, 512, dm, 256, 256, 1)
    assert np.abs(dm.to_numpy(np.float32) / g["op_mag"] - 1).max() < 2e-7
    h.close()
