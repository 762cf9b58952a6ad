"""K8 (overlap-save synchroniser) on a config-4-shaped batch generated on the device: S streams x 381 frames.
python tools/quick_os.py [S] — timing of usc_correlate_os (peaks only / with the filtered stream), oracle check on
regenerated streams, where the preamble is found."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "ultrasonic-communication_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import numpy as np, torch
import usc
from oracle import pyref as R

S = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
F, N, MB = 381, 2048, 12
SEED, LEAD, GUARD, AMP, SIGMA = 4, 40, 12, 2.0e4, 2000.0
dev = torch.device("cuda", 0)
h = usc.Handle()
st = torch.cuda.current_stream(); h.set_stream(st.cuda_stream)
pcm = torch.empty((S, F * N), dtype=torch.int32, device=dev)
h.synth_streams(SEED, 0, S, F, F * N, LEAD, MB, GUARD, AMP, SIGMA, pcm, None, None)
def timeit(fn, reps=3):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    for _ in range(reps): fn()
    e1.record(st); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
nb = S * (F - 1)
mv = torch.empty(nb, dtype=torch.float32, device=dev); mi = torch.empty(nb, dtype=torch.int32, device=dev)
ms = timeit(lambda: h.correlate_os(pcm, usc.PCM_I32, S, F, F * N, False, None, mv, mi))
gb = S * F * N * 4 / 1e9
print("K8 correlate_os (peaks): %d streams x %d frames (%.1f GB): %.2f ms  %.1f Mblocks/s  %.0f GB/s input (%.3f of 6552)" % (S, F, gb, ms, nb / ms / 1e3, gb / ms * 1e3, gb / ms * 1e3 / 6552))
g = R.arm_mult_f32(h.table("down"), h.table("hann"))
for s in (0, S - 1):
    p1, _, _ = R.synth_streams(SEED, s, 1, F, LEAD, MB, GUARD, AMP, SIGMA)
    wo, wv, wi = R.correlate_os(g, p1[0].reshape(F, N))
    assert np.array_equal(mi[s * (F - 1):(s + 1) * (F - 1)].cpu().numpy().astype(np.uint32), wi), s
    assert np.array_equal(mv[s * (F - 1):(s + 1) * (F - 1)].cpu().numpy().view(np.uint32), wv.view(np.uint32)), s
print("K8 == oracle on streams 0, %d (regenerated on the CPU)" % (S - 1))
if S <= 4096:
    out = torch.empty((nb, N), dtype=torch.float32, device=dev)
    ms = timeit(lambda: h.correlate_os(pcm, usc.PCM_I32, S, F, F * N, False, out, mv, mi))
    print("K8 correlate_os (+ filtered stream): %.2f ms  %.0f GB/s in+out" % (ms, 2 * gb / ms * 1e3))
# few long streams: segments
S2 = 64
ms = timeit(lambda: h.correlate_os(pcm, usc.PCM_I32, S2, F, F * N, False, None, mv, mi))
print("K8 correlate_os, %d streams (segmented): %.3f ms  %.0f GB/s input" % (S2, ms, S2 * F * N * 4 / 1e6 / ms))
