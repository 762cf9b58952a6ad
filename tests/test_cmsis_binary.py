"""The oracle against the reference's OWN CMSIS-DSP machine code.

tests/golden/cmsis_binary_vectors.npz was produced by tools/cmsis_emu/make_vectors.py, which links members of the
reference's vendored archive (receiver/Drivers/CMSIS/Lib/libarm_cortexM4lf_math.a, CMSIS-DSP V1.4.5b, GCC 5.4 Thumb-2)
and interprets their instructions (tools/cmsis_emu/thumb2.py): every expected value below came out of the reference's
real arm_sin_cos_f32 / arm_cos_f32 / arm_rfft_fast_f32 / arm_cfft_f32 / arm_cmplx_mult_cmplx_f32 / arm_cmplx_mag_f32 /
arm_max_f32 / arm_fir_f32 ... code, called in the order the reference's C sources call them.

Bars: tables and element-wise products that the oracle claims to restate exactly -> bit-identical; everything that
passes through an FFT (CMSIS uses a radix-8 plan, the oracle its own canonical [.,32,32] plan) -> the north-star's
1e-4 relative bound (measured: a few 1e-7), integer results (peak bins) identical.
"""
import os

import numpy as np
import pytest

from oracle import pyref as R

GOLD = os.path.join(os.path.dirname(__file__), "golden", "cmsis_binary_vectors.npz")
TOL = 1e-4                                     # BASELINE.json north_star: 1e-4 relative (fp32)


@pytest.fixture(scope="module")
def g():
    return np.load(GOLD)


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


def rel_to_peak(got, want):
    want = np.asarray(want, np.float64)
    return float(np.abs(np.asarray(got, np.float64) - want).max() / np.abs(want).max())


def test_sine_table_and_trig_functions_are_bit_identical_to_the_binary(g):
    import math
    lit = np.array([np.float32("%.8f" % math.sin(2 * math.pi * k / 512)) for k in range(513)], np.float32)
    assert np.array_equal(bits(lit), bits(g["sinTable_f32"]))              # the table as stored in the archive
    sc = np.array([R.arm_sin_cos_f32(t) for t in g["sin_cos_theta"]], np.float32)
    assert np.array_equal(bits(sc[:, 0]), bits(g["sin_cos_sin"]))
    assert np.array_equal(bits(sc[:, 1]), bits(g["sin_cos_cos"]))
    assert np.array_equal(bits(R.arm_cos_f32(g["cos_x"])), bits(g["cos_y"]))


def test_receiver_tables_are_bit_identical_to_the_binary(g):
    """up/down reference chirps through the archive's arm_sin_cos_f32 (receiver/Src/chirp.c:16-40), Hann through its
    arm_cos_f32 (receiver/Src/main.c:390-393)"""
    rx = R.RefReceiver()
    assert np.array_equal(bits(rx.table("up_chirp")), bits(g["rx_up_chirp"]))
    assert np.array_equal(bits(rx.table("down_chirp")), bits(g["rx_down_chirp"]))
    assert np.array_equal(bits(rx.table("hann")), bits(g["rx_hann"]))
    assert rx.bandwidth2 == int(g["rx_bw2"][0])
    c = R.RefCompressor()
    assert np.array_equal(bits(c.table("window")), bits(g["cc_hann"]))
    assert np.array_equal(bits(R.generate_ref_chirp("T", 2048, 100000.0, 17000.0, 18000.0, 0.0, np.float32(-3.14159265358979 / 2.0), True)),
                          bits(g["cc_chirp_up"]))
    assert np.array_equal(bits(R.generate_ref_chirp("T", 2048, 100000.0, 17000.0, 18000.0, 0.0, np.float32(-3.14159265358979 / 2.0), False)),
                          bits(g["cc_chirp_down"]))


def rx_frames(g):
    pcm = np.concatenate([R.synth_frames(int(s), int(f), int(n), a, sg)[0] for s, f, n, a, sg in g["rx_cases"]])
    assert int(np.bitwise_xor.reduce(pcm.view(np.uint32).ravel())) == int(g["rx_pcm_crc"][0])
    return pcm


def test_receiver_chain_against_the_binary(g):
    """pipeline() + arm_max_f32 (receiver/Src/main.c:163-215) for both hypotheses on 32 frames (20 of them the bench's
    config-2 dataset at -5 dB): peak bins identical, peak magnitudes and whole spectra within 1e-4"""
    rx = R.RefReceiver()
    pcm = rx_frames(g)
    mu, iu, md, idn = rx.demod_frames(pcm)
    assert np.array_equal(iu, g["rx_peak_idx"][:, 0]) and np.array_equal(idn, g["rx_peak_idx"][:, 1])
    assert np.abs(mu / g["rx_peak"][:, 0] - 1).max() < TOL and np.abs(md / g["rx_peak"][:, 1] - 1).max() < TOL
    fast = rx.demod_frames_fast(pcm)
    assert np.array_equal(fast[1], iu) and np.array_equal(fast[3], idn)
    worst = 0.0
    for f in range(len(pcm)):
        for h, up in ((0, True), (1, False)):
            m = rx.pipeline(pcm[f].astype(np.float32), up)[:1024]
            worst = max(worst, rel_to_peak(m[1:], g["rx_mag"][f, h][1:]))        # bin 0 is the packed (DC, Nyquist) pair
            assert abs(m[0] / g["rx_mag"][f, h][0] - 1) < TOL
    assert worst < TOL, worst
    assert worst < 2e-6                                                           # what it actually is (2.5e-7)


def test_compression_chain_against_the_binary(g):
    """compress_chirp (experiments/chirp_compression_time_domain/Src/chirp.c:78-83): window, RFFT, packed complex
    multiply by H_down, inverse RFFT, arm_max_f32 over the lags (main.c:189)"""
    c = R.RefCompressor()
    assert rel_to_peak(c.table("H_up"), g["cc_H_up"]) < TOL and rel_to_peak(c.table("H_down"), g["cc_H_down"]) < TOL
    s, f, n, a, sg = g["cc_case"]
    pcm, _ = R.synth_frames(int(s), int(f), int(n), a, sg, n=2048, fs=100000.0, f0=17000.0, f1=18000.0)
    for k in range(len(pcm)):
        out = c.compress(pcm[k].astype(np.float32))
        assert rel_to_peak(out, g["cc_out"][k]) < TOL
        assert int(np.argmax(out)) == int(g["cc_idx"][k])
        assert abs(out.max() / g["cc_max"][k] - 1) < TOL
    mv, mi = c.compress_frames(pcm)
    assert np.array_equal(mi, g["cc_idx"])


def test_transforms_against_the_binary(g):
    for n in (256, 1024, 4096):
        r = R.Rfft(n)
        spec = r(g["rfft%d_in" % n])
        assert rel_to_peak(spec, g["rfft%d_out" % n]) < TOL
        assert rel_to_peak(r(g["rfft%d_out" % n], inverse=True), g["rifft%d_out" % n]) < TOL
        assert rel_to_peak(g["rifft%d_out" % n], g["rfft%d_in" % n]) < TOL         # the binary's own round trip
    for n in (1024, 2048):
        cf = R.Cfft(n)
        assert rel_to_peak(cf(g["cfft%d_in" % n]), g["cfft%d_out" % n]) < TOL
        assert rel_to_peak(cf(g["cfft%d_in" % n], inverse=True), g["cifft%d_out" % n]) < TOL


def test_iq_front_end_against_the_binary(g, fir_taps):
    """carrier tables by the archive's arm_sin_cos_f32, mix, arm_fir_f32 on both rails with the state carried over
    three frames (experiments/iq_modulation/Src/iq_modem.c:34-66)"""
    assert np.array_equal(bits(fir_taps.astype(np.float32)), bits(g["iq_taps"]))
    q = R.RefIq(fir_taps.astype(np.float32)[::-1].copy())
    assert np.array_equal(bits(q.table("carrier_sin")), bits(g["iq_carrier_sin"]))
    assert np.array_equal(bits(q.table("carrier_cos")), bits(g["iq_carrier_cos"]))
    c = g["iq_case"]
    pcm, _ = R.synth_iq_frames(int(c[0]), int(c[1]), int(c[2]), c[3], c[4], int(c[5]), c[6], c[7], c[8])
    fi, fq = R.Fir(fir_taps.astype(np.float32), 2048), R.Fir(fir_taps.astype(np.float32), 2048)
    for k in range(len(pcm)):
        x = pcm[k].astype(np.float32)
        i_ = fi(R.arm_mult_f32(x, q.table("carrier_cos")))
        q_ = fq(R.arm_mult_f32(x, q.table("carrier_sin")))
        assert rel_to_peak(i_, g["iq_fir_i"][k]) < TOL and rel_to_peak(q_, g["iq_fir_q"][k]) < TOL


def test_small_operators_against_the_binary(g):
    a, b = g["op_a"], g["op_b"]
    assert np.array_equal(bits(R.arm_mult_f32(a, b)), bits(g["op_mult"]))           # one rounding: exact
    assert np.array_equal(bits(R.arm_scale_f32(a, np.float32(0.022097087))), bits(g["op_scale"]))
    assert np.array_equal(bits(R.arm_cmplx_mult_real_f32(a, b[:256])), bits(g["op_cmul_real"]))
    # the canonical complex product / magnitude fuse one multiply into an FMA; the binary rounds
    # every product: differences stay at the last bit
    assert rel_to_peak(R.arm_cmplx_mult_cmplx_f32(a, b), g["op_cmul"]) < 2e-7
    assert np.abs(R.arm_cmplx_mag_f32(a) / g["op_mag"] - 1).max() < 2e-7
    v, i = R.arm_max_f32(g["op_max_in"])
    assert v == g["op_max"][0] and i == int(g["op_max_idx"][0]) == 100              # a tie keeps the first index
    assert abs(R.arm_mean_f32(a) / g["op_mean"][0] - 1) < 1e-5
    assert abs(R.arm_mean_f32(a[:37]) / g["op_mean"][1] - 1) < 1e-5
