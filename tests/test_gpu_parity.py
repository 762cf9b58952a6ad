"""T3 (GPU): the CUDA path through the C-ABI against the CPU oracle — bit-exact for integers AND
floats (canonical arithmetic), against the device golden captures, and through size-independent
properties at the full config-2 size."""
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


def test_native_library_is_loaded(h):
    import ctypes
    assert any("libusc.so" in line for line in open("/proc/self/maps"))
    assert h.geometry() == (78, 156, 1892)
    for name, oname in (("hann", "hann"), ("up", "up_chirp"), ("down", "down_chirp")):
        assert np.array_equal(h.table(name), R.RefReceiver().table(oname))


@pytest.mark.parametrize("nframes", [1, 3, 4, 5, 127, 1024, 4099])
@pytest.mark.parametrize("dtype", [np.int32, np.float32])
def test_demod_frames_bit_exact(h, rx, nframes, dtype):
    pcm, bits = synth.make_frames(nframes, seed_noise=3 + nframes, dtype=dtype)
    mu, iu, md, idn = rx.demod_frames(pcm, nthreads=8)
    gu, giu, gd, gid, gbit = h.demod_frames_host(pcm)
    assert np.array_equal(giu, iu) and np.array_equal(gid, idn)            # integers: exact
    assert np.array_equal(gu.view(np.uint32), mu.view(np.uint32))          # floats: bit-identical
    assert np.array_equal(gd.view(np.uint32), md.view(np.uint32))
    assert np.array_equal(gbit, (~(md > mu)).astype(np.uint8))


def test_demod_empty_and_argument_errors(h):
    d = h.empty(16)
    h.demod_frames(d, usc.PCM_I32, 0)                     # empty batch is a no-op
    with pytest.raises(usc.UscError) as e:
        h.demod_frames(d, 7, 1)
    assert e.value.code == usc.USC_ERR_ARGUMENT
    with pytest.raises(usc.UscError):
        h.demod_frames(d.ptr + 4, usc.PCM_I32, 1)         # frames must be 8-byte aligned
    with pytest.raises(usc.UscError):
        h.demod_frames(None, usc.PCM_I32, 1)


@pytest.mark.parametrize("case", ["silence", "full_scale", "impulse", "dc", "low_snr"])
def test_demod_edge_inputs(h, rx, case):
    if case == "silence":
        pcm = np.zeros((8, N), np.int32)
    elif case == "full_scale":
        pcm = np.where(np.random.default_rng(1).random((8, N)) < 0.5, -(2 ** 31), 2 ** 31 - 256).astype(np.int32)
    elif case == "impulse":
        pcm = np.zeros((8, N), np.int32)
        pcm[np.arange(8), np.arange(8) * 250] = 1 << 30
    elif case == "dc":
        pcm = np.full((8, N), 12345 * 256, np.int32)
    else:
        pcm, _ = synth.make_frames(8, snr_db=-30.0)
    want = rx.demod_frames(pcm)
    got = h.demod_frames_host(pcm)
    for g, w in zip(got[:4], want):
        assert np.array_equal(g.view(np.uint32), w.view(np.uint32))
    if case == "silence":
        assert np.all(got[1] == 0) and np.all(got[0] == 0.0)   # all-equal mags: first bin wins


def test_golden_device_captures_through_gpu_ops(h, device_triples):
    """experiments/basic chain on the GPU operators: cast, Hann multiply, rfft, mag, scale vs the
    Cortex-M4 captures (1e-4 on bins >= 1 kHz above 1 % of peak; same arg-max bin)."""
    ok = device_triples["consistent"]
    raw = device_triples["raw"][ok]
    B = raw.shape[0]
    d_raw = h.buffer(raw)
    d_x = h.empty(4 * B * N)
    d_hann = h.buffer(h.table("hann"))
    h.i32_to_f32(d_raw, d_x, B * N)
    h.arm_mult_f32(d_x, N, d_hann, 0, d_x, N, N, B)
    flt = d_x.to_numpy(np.float32).reshape(B, N)
    assert np.abs(flt.astype(np.float64) - device_triples["flt"][ok]).max() <= 5.01e-7
    h.arm_rfft_fast_f32(N, d_x, d_x, 0, B)                 # in place (hazard H2 defined)
    d_mag = h.empty(4 * B * (N // 2))
    h.arm_cmplx_mag_f32(d_x, N, d_mag, N // 2, N // 2, B)
    h.arm_scale_f32(d_mag, float(np.float32(1.0) / np.sqrt(np.float32(N))), d_mag, N // 2, B)
    d_val, d_idx = h.empty(4 * B), h.empty(4 * B)
    mag = d_mag.to_numpy(np.float32).reshape(B, N // 2)
    for i in range(B):
        dev = device_triples["fft_mag"][ok][i]
        sel = device_triples["fft_freq"][ok][i] >= 1000.0
        big = sel & (dev >= 0.01 * dev[sel].max())
        assert (np.abs(mag[i][big] - dev[big]) / dev[big]).max() < 1e-4
        assert np.argmax(np.where(sel, mag[i], 0)) == np.argmax(np.where(sel, dev, 0))
    # and bit-identical to the oracle running the same operators
    rfft = R.Rfft(N)
    hann = R.hann_window(N)
    for i in range(B):
        o = R.arm_scale_f32(R.arm_cmplx_mag_f32(rfft(R.arm_mult_f32(raw[i].astype(np.float32), hann))),
                            np.float32(1.0) / np.sqrt(np.float32(N)))
        assert np.array_equal(o.view(np.uint32), mag[i].view(np.uint32))


@pytest.mark.parametrize("n", [32, 64, 128, 256, 512, 1024, 2048, 4096, 8192])
def test_rfft_operator_bit_exact_all_lengths(h, n):
    rng = np.random.default_rng(n)
    B = 5
    x = (rng.standard_normal((B, n)) * 1e4).astype(np.float32)
    d = h.buffer(x)
    o = h.empty(x.nbytes)
    h.arm_rfft_fast_f32(n, d, o, 0, B)
    got = o.to_numpy(np.float32).reshape(B, n)
    r = R.Rfft(n)
    want = np.stack([r(x[i]) for i in range(B)])
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))
    h.arm_rfft_fast_f32(n, o, o, 1, B)                      # inverse, in place
    back = o.to_numpy(np.float32).reshape(B, n)
    wantb = np.stack([r(want[i], inverse=True) for i in range(B)])
    assert np.array_equal(back.view(np.uint32), wantb.view(np.uint32))
    assert np.abs(back - x).max() <= 4e-6 * np.abs(x).max()


@pytest.mark.parametrize("n", [16, 64, 1024, 2048, 4096])
def test_cfft_operator_bit_exact(h, n):
    rng = np.random.default_rng(100 + n)
    B = 3
    x = rng.standard_normal((B, 2 * n)).astype(np.float32)
    c = R.Cfft(n)
    for inv in (False, True):
        d = h.buffer(x)
        h.arm_cfft_f32(n, d, inv, B)
        got = d.to_numpy(np.float32).reshape(B, 2 * n)
        want = np.stack([c(x[i], inverse=inv) for i in range(B)])
        assert np.array_equal(got.view(np.uint32), want.view(np.uint32))


def test_fft_argument_errors(h):
    d = h.empty(1 << 16)
    for bad in (0, 16, 48, 131072):
        with pytest.raises(usc.UscError) as e:
            h.arm_rfft_fast_f32(bad, d, d, 0, 1)
        assert e.value.code == usc.USC_ERR_ARGUMENT
    with pytest.raises(usc.UscError):
        h.arm_cfft_f32(65536, d, 0, 1)


def test_elementwise_operators_bit_exact(h):
    rng = np.random.default_rng(11)
    B, L = 7, 300                                           # ragged: not a multiple of 4 or 32
    a = rng.standard_normal((B, L)).astype(np.float32)
    b = rng.standard_normal((B, L)).astype(np.float32)
    da, db, do = h.buffer(a), h.buffer(b), h.empty(a.nbytes)
    h.arm_mult_f32(da, L, db, L, do, L, L, B)
    assert np.array_equal(do.to_numpy(np.float32).reshape(B, L), a * b)
    h.arm_mult_f32(da, L, db, 0, do, L, L, B)               # broadcast operand (stride 0)
    assert np.array_equal(do.to_numpy(np.float32).reshape(B, L), a * b[0])
    h.arm_cmplx_mult_cmplx_f32(da, L, db, 0, do, L, L // 2, B)
    want = np.stack([R.arm_cmplx_mult_cmplx_f32(a[i], b[0]) for i in range(B)])
    assert np.array_equal(do.to_numpy(np.float32).reshape(B, L).view(np.uint32), want.view(np.uint32))
    dm = h.empty(4 * B * (L // 2))
    h.arm_cmplx_mag_f32(da, L, dm, L // 2, L // 2, B)
    want = np.stack([R.arm_cmplx_mag_f32(a[i]) for i in range(B)])
    assert np.array_equal(dm.to_numpy(np.float32).reshape(B, L // 2).view(np.uint32), want.view(np.uint32))
    r = rng.standard_normal((B, L // 2)).astype(np.float32)
    dr = h.buffer(r)
    h.arm_cmplx_mult_real_f32(da, L, dr, L // 2, do, L, L // 2, B)
    want = np.stack([R.arm_cmplx_mult_real_f32(a[i], r[i]) for i in range(B)])
    assert np.array_equal(do.to_numpy(np.float32).reshape(B, L), want)
    h.arm_scale_f32(da, 0.3, do, L, B)
    assert np.array_equal(do.to_numpy(np.float32).reshape(B, L), a * np.float32(0.3))
    # max with ties and mean with a cancellation-prone order
    t = a.copy()
    t[:, 17] = t[:, 250] = 99.0
    dt, dv, di = h.buffer(t), h.empty(4 * B), h.empty(4 * B)
    h.arm_max_f32(dt, L, L, dv, di, B)
    assert np.all(di.to_numpy(np.uint32) == 17) and np.all(dv.to_numpy(np.float32) == 99.0)
    h.arm_mean_f32(dt, L, L, dv, B)
    assert np.array_equal(dv.to_numpy(np.float32), np.array([R.arm_mean_f32(t[i]) for i in range(B)], np.float32))
    # ingest cast on an unaligned length
    iv = rng.integers(-2 ** 31, 2 ** 31 - 1, size=1001, dtype=np.int64).astype(np.int32)
    dvv, dff = h.buffer(iv), h.empty(4 * 1001)
    h.i32_to_f32(dvv, dff, 1001)
    assert np.array_equal(dff.to_numpy(np.float32), iv.astype(np.float32))


def test_fir_operator_bit_exact(h, fir_taps):
    rng = np.random.default_rng(12)
    B, L = 5, 2048
    x = (rng.standard_normal((2, B, L)) * 1e3).astype(np.float32)
    coeffs = fir_taps.astype(np.float32)[::-1].copy()
    state = h.buffer(np.zeros((B, 26), np.float32))
    firs = [R.Fir(coeffs, L) for _ in range(B)]
    for blk in range(2):
        d, o = h.buffer(x[blk]), h.empty(x[blk].nbytes)
        h.arm_fir_f32(coeffs, state, d, o, L, B)
        got = o.to_numpy(np.float32).reshape(B, L)
        want = np.stack([firs[i](x[blk, i]) for i in range(B)])
        assert np.array_equal(got.view(np.uint32), want.view(np.uint32))


def test_pipeline_and_dsp_bit_exact(h, rx):
    pcm, _ = synth.make_frames(64, seed_noise=21, dtype=np.float32)
    d, o = h.buffer(pcm), h.empty(pcm.nbytes)
    for updown in (usc.UP, usc.DOWN):
        h.pipeline(d, o, updown, 64)
        got = o.to_numpy(np.float32).reshape(64, N)
        want = np.stack([rx.pipeline(pcm[i], up=bool(updown)) for i in range(64)])
        assert np.array_equal(got.view(np.uint32), want.view(np.uint32))
    # dsp(): 21 streams with 3-frame fifos and arbitrary (odd, edge) sync positions
    S = 21
    fifo = pcm[:3 * S].reshape(S, 3 * N).copy()
    pos = np.array([0, 1, 255, 256, 1023, 1024, 1280, 2047, 2048, 2049, 3000, 4095, 4096, 7, 512, 768, 1536, 1792, 2304, 3333, 4000], np.uint32)
    mean = np.linspace(1e7, 5e8, S).astype(np.float32)
    d_f, d_p, d_m = h.buffer(fifo), h.buffer(pos), h.buffer(mean)
    d_h = h.empty(48 * S)
    for updown in (usc.UP, usc.DOWN):
        h.dsp(d_f, 3 * N, d_p, d_m, updown, d_h, S)
        got = d_h.to_numpy(usc.history_dtype)
        for s in range(S):
            w = rx.dsp(fifo[s], int(pos[s]), float(mean[s]), up=bool(updown))
            g = got[s]
            assert (g["max_idx"], g["max_idx_left"], g["max_idx_right"]) == (w.max_idx, w.max_idx_left, w.max_idx_right)
            assert (g["max_freq"], g["max_freq_left"], g["max_freq_right"]) == (w.max_freq, w.max_freq_left, w.max_freq_right)
            for k in ("mag_max", "mag_max_left", "mag_max_right", "mag_mean", "snr"):
                assert np.float32(g[k]).view(np.uint32) == np.float32(getattr(w, k)).view(np.uint32), (s, k)


def test_compress_chirp_bit_exact():
    cfg = usc.default_config(fs=100000.0, f0=17000.0, f1=18000.0, chirp_variant=usc.CHIRP_T, window=usc.HANN_SYMMETRIC)
    hc = usc.Handle(cfg)
    c = R.RefCompressor()
    assert np.array_equal(hc.table("hann"), c.table("window"))
    assert np.array_equal(hc.table("H_down").view(np.uint32), c.table("H_down").view(np.uint32))
    assert np.array_equal(hc.table("H_up").view(np.uint32), c.table("H_up").view(np.uint32))
    rng = np.random.default_rng(31)
    p = hc.table("up")
    F = 37
    pcm = np.stack([np.rint(np.roll(p, int(rng.integers(0, N))) * 20000 + rng.standard_normal(N) * 8000) for _ in range(F)])
    pcm = (pcm.astype(np.int64) * 256).astype(np.int32)
    d = hc.buffer(pcm)
    d_out, d_v, d_i = hc.empty(4 * F * N), hc.empty(4 * F), hc.empty(4 * F)
    for use_up in (False, True):
        hc.compress_chirp(d, usc.PCM_I32, F, use_up, d_out, d_v, d_i)
        wv, wi = c.compress_frames(pcm, use_up=use_up, nthreads=4)
        assert np.array_equal(d_i.to_numpy(np.uint32), wi)
        assert np.array_equal(d_v.to_numpy(np.float32).view(np.uint32), wv.view(np.uint32))
        out = d_out.to_numpy(np.float32).reshape(F, N)
        want = np.stack([c.compress(pcm[i].astype(np.float32), use_up) for i in range(F)])
        assert np.array_equal(out.view(np.uint32), want.view(np.uint32))
    hc.compress_chirp(d, usc.PCM_I32, F, False, None, d_v, d_i)      # peaks only
    assert np.array_equal(d_i.to_numpy(np.uint32), c.compress_frames(pcm)[1])
    hc.close()


def test_full_size_properties(h, rx):
    """Config 2 at full size (4096 streams x 38 frames = 155648 frames, 1.275 GB): properties that
    need no full-size oracle run — batch-split invariance, permutation equivariance, a sampled
    oracle check, and the bit decisions against the transmitted bits."""
    import torch
    F = 4096 * 38
    base, bits = synth.make_frames(4096, snr_db=-5.0)
    tb = torch.from_numpy(base).cuda()
    perm = torch.randperm(4096, generator=torch.Generator().manual_seed(5))
    big = torch.empty((F, N), dtype=torch.int32, device="cuda")
    for r in range(38):
        big[r * 4096:(r + 1) * 4096] = tb if r % 2 == 0 else tb[perm.cuda()]
    outs = [torch.empty(F, dtype=torch.float32, device="cuda"), torch.empty(F, dtype=torch.int32, device="cuda"),
            torch.empty(F, dtype=torch.float32, device="cuda"), torch.empty(F, dtype=torch.int32, device="cuda"),
            torch.empty(F, dtype=torch.uint8, device="cuda")]
    torch.cuda.synchronize()
    h.demod_frames(big, usc.PCM_I32, F, *outs)
    h.sync()
    torch.cuda.synchronize()
    mu, iu, md, idn, bit = [o.cpu().numpy() for o in outs]
    # every repetition block equals block 0 (even) or its permutation (odd): frames are independent
    for r in range(38):
        sl = slice(r * 4096, (r + 1) * 4096)
        ref = slice(0, 4096)
        if r % 2 == 0:
            assert np.array_equal(iu[sl], iu[ref]) and np.array_equal(mu[sl], mu[ref])
        else:
            assert np.array_equal(iu[sl], iu[ref][perm.numpy()]) and np.array_equal(md[sl], md[ref][perm.numpy()])
    # block 0 against the oracle, bit-exact
    wmu, wiu, wmd, wid = rx.demod_frames(base, nthreads=8)
    assert np.array_equal(iu[:4096].astype(np.uint32), wiu) and np.array_equal(idn[:4096].astype(np.uint32), wid)
    assert np.array_equal(mu[:4096].view(np.uint32), wmu.view(np.uint32))
    assert np.array_equal(md[:4096].view(np.uint32), wmd.view(np.uint32))
    assert (bit[:4096] == bits).mean() > 0.9


def test_spectrum_analyzer_reproduces_device_captures(device_triples):
    """experiments/basic fft() end to end on the GPU (usc_spectrum_analyzer) vs the Cortex-M4 captures:
    every printed magnitude (AC-coupled bins = 1.0 included) within 1e-4 relative (+1e-6 print
    resolution) and the same peak bin, at the three sampling rates the captures were taken with
    (fs = 80 MHz / divider / 32 in integer arithmetic, experiments/basic*/Src/main.c:240-243)."""
    ok = device_triples["consistent"]
    names = device_triples["names"][ok]
    raw, dev_mag, freq = device_triples["raw"][ok], device_triples["fft_mag"][ok], device_triples["fft_freq"][ok]
    checked = 0
    for fs in (100000.0, 48076.0, 41666.0):
        sel = np.array([abs(f[1000] * 2048 / 1000 - fs) < 30 for f in freq])
        if not sel.any():
            continue
        h = usc.Handle(usc.default_config(fs=fs))
        B = int(sel.sum())
        d = h.buffer(raw[sel])
        d_mag, d_db, d_pk, d_pi = h.empty(4 * B * 1024), h.empty(4 * B * 1024), h.empty(4 * B), h.empty(4 * B)
        h.spectrum_analyzer(d, usc.PCM_I32, B, 1000.0, d_mag, d_db, d_pk, d_pi)
        h.sync()
        mag = d_mag.to_numpy(np.float32).reshape(B, 1024)
        db = d_db.to_numpy(np.float32).reshape(B, 1024)
        pi = d_pi.to_numpy(np.uint32)
        want = dev_mag[sel]
        fr = freq[sel]
        for i in range(B):
            ac = fr[i] < 1000.0 - 0.05                      # printed with one decimal
            edge = np.abs(fr[i] - 1000.0) <= 0.05
            assert np.all(mag[i][ac] == 1.0)
            body = ~ac & ~edge
            big = body & (want[i] >= 0.01 * want[i][body].max())
            assert (np.abs(mag[i][big] - want[i][big]) / want[i][big]).max() < 1e-4, names[sel][i]
            assert np.abs(mag[i][body] - want[i][body]).max() <= 1e-4 * want[i][body].max() + 1e-6
            assert int(pi[i]) == int(np.argmax(np.where(edge, 0, want[i])))
            assert np.abs(db[i] - 10.0 * np.log10(mag[i].astype(np.float64))).max() < 1e-4
            checked += 1
        h.close()
    assert checked == 23


@pytest.mark.parametrize("nframes", [1, 2, 3, 8, 255, 4099])
@pytest.mark.parametrize("dtype", [np.int32, np.float32])
def test_single_hypothesis_pair_kernel_bit_exact(h, rx, nframes, dtype):
    """usc_demod_frames with only one hypothesis' outputs requested = dsp(.., UP) or dsp(.., DOWN) alone
    (receiver/Src/main.c:183-215); frames go through the packed core two at a time (odd counts too)."""
    pcm, _ = synth.make_frames(nframes, seed_noise=77 + nframes, dtype=dtype)
    want = rx.demod_frames(pcm, nthreads=8)
    fmt = usc.PCM_I32 if dtype == np.int32 else usc.PCM_F32
    d = h.buffer(pcm)
    m, i = h.empty(4 * nframes), h.empty(4 * nframes)
    h.demod_frames(d, fmt, nframes, mag_up=m, idx_up=i)
    h.sync()
    assert np.array_equal(m.to_numpy(np.float32).view(np.uint32), want[0].view(np.uint32))
    assert np.array_equal(i.to_numpy(np.uint32), want[1])
    h.demod_frames(d, fmt, nframes, mag_down=m, idx_down=i)
    h.sync()
    assert np.array_equal(m.to_numpy(np.float32).view(np.uint32), want[2].view(np.uint32))
    assert np.array_equal(i.to_numpy(np.uint32), want[3])
    h.demod_frames(d, fmt, nframes, idx_up=i)                   # indices only
    h.sync()
    assert np.array_equal(i.to_numpy(np.uint32), want[1])


def test_host_buffer_path_matches_device_path(h, rx):
    """usc_demod_frames_host (chunked H2D / K1 / D2H through three stream lanes) on pageable numpy memory,
    with a chunk size that does not divide the frame count."""
    nframes = 1000
    pcm, _ = synth.make_frames(nframes, seed_noise=5)
    want = rx.demod_frames(pcm, nthreads=8)
    h.host_workspace(96)
    mu, md = np.empty(nframes, np.float32), np.empty(nframes, np.float32)
    iu, idn = np.empty(nframes, np.uint32), np.empty(nframes, np.uint32)
    bit = np.empty(nframes, np.uint8)
    h.demod_frames_hostbuf(pcm, usc.PCM_I32, nframes, mu, iu, md, idn, bit)
    for g, w in zip((mu, iu, md, idn), want):
        assert np.array_equal(g.view(np.uint32), w.view(np.uint32))
    assert np.array_equal(bit, (~(want[2] > want[0])).astype(np.uint8))
    h.host_workspace(4096)


def test_scan4_freq_domain_experiment_bit_exact():
    """experiments/chirp_compression_freq_domain: 4-offset scan (offsets 0, N/4, N/2, 3N/4 over a 2N buffer),
    real down-chirp (variant F), right window on the magnitudes, left window on the leftover packed
    spectrum floats of the in-place buffer."""
    cfg = usc.default_config(fs=100000.0, f0=17000.0, f1=18000.0, chirp_variant=usc.CHIRP_F)
    hf = usc.Handle(cfg)
    assert hf.geometry()[0] == 20                                # (F2-F1)*N/fs = 20.48 -> 20
    rng = np.random.default_rng(17)
    B = 9
    up = R.generate_ref_chirp("F", N, 100000.0, 17000.0, 18000.0, 0.0, 0.0, 1)
    bufs = np.zeros((B, 2 * N), np.float32)
    for b in range(B):
        off = int(rng.integers(0, N))
        bufs[b, off:off + N] += up * 20000
        bufs[b] += rng.standard_normal(2 * N).astype(np.float32) * 3000
    d, o = hf.buffer(bufs), hf.empty(B * 4 * 16)
    hf.scan4(d, B, o)
    hf.sync()
    got = o.to_numpy(usc.scan_entry_dtype).reshape(B, 4)
    for b in range(B):
        want = R.scan4(bufs[b])
        for i in range(4):
            g = got[b, i]
            assert (g["max_idx_right"], g["max_idx_left"]) == (want[i][1], want[i][3])
            assert np.float32(g["mag_max_right"]).view(np.uint32) == np.float32(want[i][0]).view(np.uint32)
            assert np.float32(g["mag_max_left"]).view(np.uint32) == np.float32(want[i][2]).view(np.uint32)
    hf.close()


def test_two_devices_in_one_process():
    """Handles on different GPUs of one process: per-device kernel configuration and device switching inside
    the library (skipped on a one-GPU box; the bench's multi-GPU mode is one process per GPU)."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    pcm, _ = synth.make_frames(64)
    want = R.RefReceiver().demod_frames(pcm, nthreads=4)
    hs = [usc.Handle(device=d) for d in (0, 1)]
    outs = [h.demod_frames_host(pcm) for h in hs]
    outs += [h.demod_frames_host(pcm) for h in reversed(hs)]
    for got in outs:
        for g, w in zip(got[:4], want):
            assert np.array_equal(g.view(np.uint32), w.view(np.uint32))
    for h in hs:
        h.close()


@pytest.mark.parametrize("B", [1, 2, 3, 64, 517])
def test_warp_fft_operators_bit_exact(B, monkeypatch):
    """The receiver's own lengths (rfft 2048, cfft 1024) run on the packed register core (k_fft_warp.cu): odd and
    even batches, out of place and in place, forward and inverse, on a handle configured for another frame
    length (tables built on first use) — equal to the oracle and to the generic shared-memory kernel."""
    rng = np.random.default_rng(B)
    hh = usc.Handle(usc.default_config(n=4096))
    x = (rng.standard_normal((B, 2048)) * 3e4).astype(np.float32)
    x[0, :] = 0.0
    r, c = R.Rfft(2048), R.Cfft(1024)
    d, o = hh.buffer(x), hh.empty(x.nbytes)
    hh.arm_rfft_fast_f32(2048, d, o, 0, B)
    fwd = o.to_numpy(np.float32).reshape(B, 2048)
    want = np.stack([r(x[i]) for i in range(B)])
    assert np.array_equal(fwd.view(np.uint32), want.view(np.uint32))
    hh.arm_rfft_fast_f32(2048, o, o, 1, B)                  # inverse in place
    wantb = np.stack([r(want[i], inverse=True) for i in range(B)])
    assert np.array_equal(o.to_numpy(np.float32).reshape(B, 2048).view(np.uint32), wantb.view(np.uint32))
    for inv in (False, True):
        dc = hh.buffer(x)
        hh.arm_cfft_f32(1024, dc, inv, B)
        wantc = np.stack([c(x[i], inverse=inv) for i in range(B)])
        assert np.array_equal(dc.to_numpy(np.float32).reshape(B, 2048).view(np.uint32), wantc.view(np.uint32))
    y = (rng.standard_normal((B, 4096)) * 3e4).astype(np.float32)           # arm_cfft_sR_f32_len2048
    c2 = R.Cfft(2048)
    for inv in (False, True):
        dc = hh.buffer(y)
        hh.arm_cfft_f32(2048, dc, inv, B)
        wantc = np.stack([c2(y[i], inverse=inv) for i in range(B)])
        assert np.array_equal(dc.to_numpy(np.float32).reshape(B, 4096).view(np.uint32), wantc.view(np.uint32))
    # the generic kernel gives the same bits (USC_FFT_GENERIC routes around the warp-level operators)
    monkeypatch.setenv("USC_FFT_GENERIC", "1")
    hh.arm_rfft_fast_f32(2048, d, o, 0, B)
    assert np.array_equal(o.to_numpy(np.float32).reshape(B, 2048).view(np.uint32), want.view(np.uint32))
    hh.close()


def test_streaming_operators_row_form_bit_exact(h):
    """mult / cmplx_mult_cmplx / cmplx_mag on aligned rows of the receiver's length take the float4 row kernels:
    broadcast tables (stride 0), in place, out of place — same bits as the oracle's scalar loops."""
    rng = np.random.default_rng(77)
    B, L = 9, 2048
    a = (rng.standard_normal((B, L)) * 1e3).astype(np.float32)
    w = rng.standard_normal(L).astype(np.float32)
    b = rng.standard_normal((B, L)).astype(np.float32)
    da, dw, db, do = h.buffer(a), h.buffer(w), h.buffer(b), h.empty(a.nbytes)
    h.arm_mult_f32(da, L, dw, 0, do, L, L, B)
    assert np.array_equal(do.to_numpy(np.float32).reshape(B, L), a * w[None, :])
    h.arm_mult_f32(da, L, db, L, do, L, L, B)
    assert np.array_equal(do.to_numpy(np.float32).reshape(B, L), a * b)
    h.arm_cmplx_mult_cmplx_f32(da, L, dw, 0, do, L, L // 2, B)
    want = np.stack([R.arm_cmplx_mult_cmplx_f32(a[i], w) for i in range(B)])
    assert np.array_equal(do.to_numpy(np.float32).reshape(B, L).view(np.uint32), want.view(np.uint32))
    dm = h.empty(4 * B * (L // 2))
    h.arm_cmplx_mag_f32(da, L, dm, L // 2, L // 2, B)
    wantm = np.stack([R.arm_cmplx_mag_f32(a[i]) for i in range(B)])
    assert np.array_equal(dm.to_numpy(np.float32).reshape(B, L // 2).view(np.uint32), wantm.view(np.uint32))
    d2 = h.buffer(a)
    h.arm_mult_f32(d2, L, dw, 0, d2, L, L, B)                                  # in place
    assert np.array_equal(d2.to_numpy(np.float32).reshape(B, L), a * w[None, :])
    d3 = h.buffer(a)
    h.arm_cmplx_mult_cmplx_f32(d3, L, db, L, d3, L, L // 2, B)                  # in place
    want = np.stack([R.arm_cmplx_mult_cmplx_f32(a[i], b[i]) for i in range(B)])
    assert np.array_equal(d3.to_numpy(np.float32).reshape(B, L).view(np.uint32), want.view(np.uint32))


def test_cmplx_mag_in_place_is_the_reference_call_form(h):
    """receiver/Src/main.c:178 calls arm_cmplx_mag_f32(signal, signal, NN): dst == src.  Magnitude e lands on an
    input of magnitude e/2, so the batched operator must order reads before writes (cross-thread race)."""
    rng = np.random.default_rng(31)
    for B, L in ((1, 2048), (37, 2048), (5, 300), (3, 65536)):
        a = rng.standard_normal((B, L)).astype(np.float32)
        want = np.stack([R.arm_cmplx_mag_f32(a[i]) for i in range(B)])
        d = h.buffer(a)
        h.arm_cmplx_mag_f32(d, L, d, L, L // 2, B)
        got = d.to_numpy(np.float32).reshape(B, L)
        assert np.array_equal(got[:, :L // 2].view(np.uint32), want.view(np.uint32)), (B, L)
        assert np.array_equal(got[:, L // 2:], a[:, L // 2:])          # the upper half of each row is not touched
    # an overlap that is not the in-place form has no defined result: refused
    a = rng.standard_normal((4, 512)).astype(np.float32)
    d = h.buffer(a)
    with pytest.raises(usc.UscError) as ei:
        h.arm_cmplx_mag_f32(d, 512, d, 256, 256, 4)
    assert ei.value.code == usc.USC_ERR_ARGUMENT


def test_pipeline_on_the_longest_frame():
    """usc_pipeline on a 65536-point handle (the tail kernel asked for n*4 bytes of shared memory)."""
    n = 65536
    hh = usc.Handle(usc.default_config(n=n))
    rx = R.RefReceiver(n=n)
    rng = np.random.default_rng(32)
    x = (rng.standard_normal((2, n)) * 1e4).astype(np.float32)
    d, o = hh.buffer(x), hh.empty(x.nbytes)
    hh.pipeline(d, o, usc.UP, 2)
    got = o.to_numpy(np.float32).reshape(2, n)
    want = np.stack([rx.pipeline(x[i], up=True) for i in range(2)])
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))
    hh.close()
