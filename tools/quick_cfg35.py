"""Throughput of the config-3 (I/Q) and config-5 (long frames) paths (operator chains)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "ultrasonic-communication_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import numpy as np, torch
import usc

dev = torch.device("cuda", 0)
st = torch.cuda.current_stream()
def timeit(fn, reps=3):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    for _ in range(reps): fn()
    e1.record(st); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
# config 3: I/Q, 16384 streams x 38 frames (use 4096 streams)
taps = np.load(os.path.join(ROOT, "tests/golden/fir_taps.npz"))["taps"].astype(np.float32)[::-1].copy()
h = usc.Handle(); h.set_stream(st.cuda_stream); h.iq_init(18000.0, 3000.0, taps, 32)
S, F, N = 4096, 38, 2048
pcm = torch.empty((S * F, N), dtype=torch.int32, device=dev)
h.synth_frames(1, 0, S * F, 2.0e4, 2.0e4, pcm)
o = [torch.empty(S * F, dtype=torch.float32, device=dev) for _ in range(2)] + [torch.empty(S * F, dtype=torch.int32, device=dev) for _ in range(2)]
b = torch.empty(S * F, dtype=torch.uint8, device=dev)
ms = timeit(lambda: h.iq_demod(pcm, usc.PCM_I32, S, F, F * N, o[0], o[2], o[1], o[3], b))
print("K5 I/Q: %d frames %.2f ms  %.1f Mframes/s  %.0f GB/s (%.1f%% of 6552)" % (S * F, ms, S * F / ms / 1e3, S * F * 8216 / ms / 1e6, S * F * 8216 / ms / 1e6 / 65.52))
h.close()
for n in (4096, 8192, 16384, 32768, 65536):
    hh = usc.Handle(usc.default_config(n=n)); hh.set_stream(st.cuda_stream)
    nf = (1 << 28) // n                       # 1 GiB of PCM
    x = torch.empty((nf, n), dtype=torch.int32, device=dev)
    hh.synth_frames(2, 0, nf, 2.0e4, 1.0e5, x)
    oo = [torch.empty(nf, dtype=torch.float32, device=dev) for _ in range(2)] + [torch.empty(nf, dtype=torch.int32, device=dev) for _ in range(2)]
    bb = torch.empty(nf, dtype=torch.uint8, device=dev)
    ms = timeit(lambda: hh.demod_frames(x, usc.PCM_I32, nf, oo[0], oo[2], oo[1], oo[3], bb))
    print("K6 n=%d: %d frames %.2f ms  %.3f Mframes/s  %.0f GB/s (%.1f%% of 6552)" % (n, nf, ms, nf / ms / 1e3, nf * (4 * n + 16) / ms / 1e6, nf * (4 * n + 16) / ms / 1e6 / 65.52))
    hh.close()
