"""Small workload touching every fused kernel once, for compute-sanitizer (memcheck / racecheck / initcheck)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "ultrasonic-communication_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import numpy as np
import usc, synth

N = 2048
h = usc.Handle()
pcm, _ = synth.make_frames(37)
out = h.demod_frames_host(pcm)                                   # K1 dual
d = h.buffer(pcm); m, i = h.empty(4 * 37), h.empty(4 * 37)
h.demod_frames(d, usc.PCM_I32, 37, mag_up=m, idx_up=i); h.sync()  # K1 pair
st = synth.make_stream(b"Hi", snr_db=20.0, nframes=70)[None]
print(h.receiver_run_host(np.repeat(st, 3, 0))[0])                # K7
dm, di = h.empty(4 * 70 * 4), h.empty(4 * 70 * 4)
ds = h.buffer(st)
for K in (1, 3):
    h.sync_search(ds, usc.PCM_I32, 1, 70, 70 * N, K, dm, di); h.sync()   # K4
f = h.buffer(pcm[:3].astype(np.float32).reshape(1, -1)); pos = h.buffer(np.array([300], np.uint32)); mean = h.buffer(np.array([1e8], np.float32))
hh = h.empty(48); h.dsp(f, 3 * N, pos, mean, usc.UP, hh, 1); h.sync()     # k_dsp2048
hc = usc.Handle(usc.default_config(fs=100000.0, f0=17000.0, f1=18000.0, chirp_variant=usc.CHIRP_T, window=usc.HANN_SYMMETRIC))
dc = hc.buffer(pcm); o = hc.empty(4 * 37 * N)
hc.compress_chirp(dc, usc.PCM_I32, 37, False, o, m, i); hc.sync()  # K2
hs = usc.Handle(usc.default_config(chirp_variant=usc.CHIRP_S))
hs.dsp(f, 3 * N, pos, mean, usc.DOWN, hh, 1); hs.sync()            # K3
x = h.buffer(np.random.default_rng(0).standard_normal((2, 65536)).astype(np.float32)); y = h.empty(2 * 65536 * 4)
h.arm_rfft_fast_f32(65536, x, y, 0, 2); h.arm_rfft_fast_f32(4096, x, y, 0, 2); h.arm_rfft_fast_f32(4096, y, y, 1, 2); h.sync()
for n, nf in ((4096, 5), (8192, 3), (16384, 2), (32768, 70), (65536, 40)):                # K6: fused long frames, cluster kernel at 65536
    hl = usc.Handle(usc.default_config(n=n))
    pl, _ = synth.make_frames(nf, n=n, seed_noise=n)
    hl.demod_frames_host(pl); hl.close()
taps = np.load(os.path.join(ROOT, "tests/golden/fir_taps.npz"))["taps"].astype(np.float32)[::-1].copy()
hq = usc.Handle(); hq.iq_init(18000.0, 3000.0, taps, 32)           # K5: whole I/Q path
pq = np.stack([synth.make_iq_stream(5, seed_bits=s)[0] for s in range(3)])
dq = hq.buffer(pq); oq = [hq.empty(4 * 15) for _ in range(4)]; bq = hq.empty(15)
hq.iq_demod(dq, usc.PCM_I32, 3, 5, 5 * N, oq[0], oq[1], oq[2], oq[3], bq); hq.sync()
xr = h.buffer(np.random.default_rng(1).standard_normal((5, 2048)).astype(np.float32)); yr = h.empty(5 * 2048 * 4)
h.arm_rfft_fast_f32(2048, xr, yr, 0, 5); h.arm_rfft_fast_f32(2048, yr, yr, 1, 5)            # warp-level FFT operators
h.arm_cfft_f32(1024, xr, 0, 5); h.arm_cfft_f32(1024, xr, 1, 5); h.arm_cfft_f32(2048, xr, 0, 2); h.arm_cfft_f32(2048, xr, 1, 2); h.sync()
ps, _ = synth.make_onoff_stream(b"A"); dps = h.buffer(ps); F1 = ps.shape[0]
lv, cd, ch, nc = h.empty(F1), h.empty(F1), h.empty(8), h.empty(4)
h.onoff_detect(dps, usc.PCM_I32, 1, F1, None, None, lv, ch, 8, nc); h.fsk_detect(dps, usc.PCM_I32, 1, F1, None, cd, None, None, ch, 8, nc); h.sync()   # legacy detectors
tr = h.buffer(np.random.default_rng(2).integers(-2000, 2000, 5000).astype(np.int16)); to = h.empty(4 * 8858)
h.resample_i16_to_pcm(tr, 5000, 3125, 1764, to, 8858); h.sync()                             # resampler
stt = h.empty(160 * 3); usc.load().usc_memset(h._h, __import__("ctypes").c_void_p(stt.ptr), 0, __import__("ctypes").c_size_t(480))
st3 = np.repeat(st, 3, 0); u3, r3 = h.empty(3 * 32), h.empty(3 * 32)
h.receiver_run_chunk(h.buffer(st3[:, :30]), usc.PCM_I32, 3, 30, 30 * N, 0, stt, u3, 32, r3)
h.receiver_run_chunk(h.buffer(np.ascontiguousarray(st3[:, 28:70])), usc.PCM_I32, 3, 40, 42 * N, 2, stt, u3, 32, r3); h.sync()   # K7 in chunks
sp = h.empty(4 * 4 * 50 * N); h.synth_streams(1, 0, 4, 50, 50 * N, 5, 2, 3, 2.0e4, 2000.0, sp); h.sync()   # stream generator
# K8 overlap-save synchroniser: several streams, segment splits, with and without the filtered stream
po, _ = synth.make_frames(3 * 9); dpo = h.buffer(po.reshape(3, -1)); oo = h.empty(4 * 3 * 8 * N); mo, io = h.empty(4 * 24), h.empty(4 * 24)
h.correlate_os(dpo, usc.PCM_I32, 3, 9, 9 * N, False, oo, mo, io); h.correlate_os(dpo, usc.PCM_I32, 3, 9, 9 * N, True, None, mo, io); h.sync()
# K5 with a filter length other than the reference's 27 taps (run-time tap count form) and the I/Q generator
hq2 = usc.Handle(); hq2.iq_init(18000.0, 3000.0, taps[:21].copy(), 32)
hq2.iq_demod(dq, usc.PCM_I32, 3, 5, 5 * N, oq[0], oq[1], oq[2], oq[3], bq); hq2.sync()
gi = h.empty(4 * 4 * N); h.synth_iq_frames(3, 0, 4, 18000.0, 3000.0, -1, 0.0, 2.0e4, 1.0e4, gi); h.sync()
# host-buffer entry points (chunked through three stream lanes)
import torch
hp = torch.from_numpy(np.ascontiguousarray(pcm)).pin_memory()
ho = [torch.empty(37, dtype=torch.float32).pin_memory() for _ in range(2)] + [torch.empty(37, dtype=torch.int32).pin_memory() for _ in range(2)]
hb = torch.empty(37, dtype=torch.uint8).pin_memory()
h.host_workspace(16); h.demod_frames_hostbuf(hp, usc.PCM_I32, 37, ho[0], ho[2], ho[1], ho[3], hb)
print("sanitize workload done")
