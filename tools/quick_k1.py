"""A/B loop for the dual-hypothesis K1: parity on 2048 frames + device-resident timing on config 2.
usage: python tools/quick_k1.py [lib.so ...]   (each library in its own process via USC_LIB; no argument = the in-tree build)"""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if len(sys.argv) > 1 and sys.argv[1] != "--child":
    for lib in sys.argv[1:]:
        env = dict(os.environ, USC_LIB=os.path.abspath(lib))
        r = subprocess.run([sys.executable, __file__, "--child"], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        print("%-40s %s" % (os.path.basename(lib), r.stdout.strip().splitlines()[-1] if r.stdout.strip() else "no output"), flush=True)
    sys.exit(0)
for p in (ROOT, os.path.join(ROOT, "ultrasonic-communication_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import numpy as np, torch
import usc, synth, bench
from oracle import pyref as R
h = usc.Handle()
pcm, bits = synth.make_frames(2048)
want = R.RefReceiver().demod_frames(pcm, nthreads=8)
got = h.demod_frames_host(pcm)
ok = all(np.array_equal(g.view(np.uint32), w.view(np.uint32)) for g, w in zip(got[:4], want))
dev = torch.device("cuda", 0)
big, _ = bench.make_device_frames(torch, h, bench.NFRAMES, dev, 0)
F = bench.NFRAMES
o = [torch.empty(F, dtype=torch.float32, device=dev) for _ in range(4)]
b = torch.empty(F, dtype=torch.uint8, device=dev)
st = torch.cuda.current_stream(); h.set_stream(st.cuda_stream)
def timeit(fn, reps=50):
    for _ in range(10): fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record(st)
    for _ in range(reps): fn()
    e1.record(st); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
ms = min(timeit(lambda: h.demod_frames(big, usc.PCM_I32, F, o[0], o[1], o[2], o[3], b)) for _ in range(3))
print("parity %s  K1 dual %.4f ms  %.1f Msym/s  %.1f%% of 6552 GB/s" % (ok, ms, F / ms / 1e3, F * 8208 / ms / 1e6 / 65.52))
