"""Quick GPU iteration loop: K1/K2 parity on a small batch + device-resident timing on config 2."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "ultrasonic-communication_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import numpy as np, torch
import usc, synth, bench
from oracle import pyref as R

h = usc.Handle()
pcm, bits = synth.make_frames(2048)
want = R.RefReceiver().demod_frames(pcm, nthreads=8)
got = h.demod_frames_host(pcm)
print("K1 parity:", all(np.array_equal(g.view(np.uint32), w.view(np.uint32)) for g, w in zip(got[:4], want)))
dev = torch.device("cuda", 0)
big, _ = bench.make_device_frames(torch, h, bench.NFRAMES, dev, 0)
F = bench.NFRAMES
o = [torch.empty(F, dtype=torch.float32, device=dev) for _ in range(4)]
b = torch.empty(F, dtype=torch.uint8, device=dev)
st = torch.cuda.current_stream(); h.set_stream(st.cuda_stream)
def timeit(fn, reps=30):
    for _ in range(5): fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record(st)
    for _ in range(reps): fn()
    e1.record(st); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
ms = timeit(lambda: h.demod_frames(big, usc.PCM_I32, F, o[0], o[1], o[2], o[3], b))
print("K1 dual   : %.3f ms  %.1f Msym/s  %.0f GB/s (%.1f%% of 6552)" % (ms, F / ms / 1e3, F * 8208 / ms / 1e6, F * 8208 / ms / 1e6 / 65.52))
# single hypothesis (pair-mode kernel)
for nf in (1, 2, 7, 2048):
    d = h.buffer(pcm[:nf]); a, b2 = h.empty(4 * nf), h.empty(4 * nf)
    ok = True
    for up in (True, False):
        if up: h.demod_frames(d, usc.PCM_I32, nf, mag_up=a, idx_up=b2)
        else: h.demod_frames(d, usc.PCM_I32, nf, mag_down=a, idx_down=b2)
        h.sync()
        wm, wi = (want[0], want[1]) if up else (want[2], want[3])
        ok &= np.array_equal(a.to_numpy(np.float32).view(np.uint32), wm[:nf].view(np.uint32)) and np.array_equal(b2.to_numpy(np.uint32), wi[:nf])
    print("K1 single parity nf=%d:" % nf, ok)
ms = timeit(lambda: h.demod_frames(big, usc.PCM_I32, F, o[0], o[1]))
print("K1 single : %.3f ms  %.1f Mframes/s  %.0f GB/s (%.1f%% of 6552)" % (ms, F / ms / 1e3, F * 8200 / ms / 1e6, F * 8200 / ms / 1e6 / 65.52))
if "--compress" in sys.argv:
    hc = usc.Handle(usc.default_config(fs=100000.0, f0=17000.0, f1=18000.0, chirp_variant=usc.CHIRP_T, window=usc.HANN_SYMMETRIC))
    hc.set_stream(st.cuda_stream)
    c = R.RefCompressor()
    d = hc.buffer(pcm[:64]); dv, di = hc.empty(4 * 64), hc.empty(4 * 64)
    hc.compress_chirp(d, usc.PCM_I32, 64, False, None, dv, di); hc.sync()
    wv, wi = c.compress_frames(pcm[:64])
    print("K2 parity:", np.array_equal(di.to_numpy(np.uint32), wi) and np.array_equal(dv.to_numpy(np.float32).view(np.uint32), wv.view(np.uint32)))
    ms = timeit(lambda: hc.compress_chirp(big, usc.PCM_I32, F, False, None, o[0], o[1]))
    print("K2 compress: %.3f ms  %.1f Mframes/s  %.0f GB/s (%.1f%% of 6552)" % (ms, F / ms / 1e3, F * 8200 / ms / 1e6, F * 8200 / ms / 1e6 / 65.52))
