"""Timing of K7 (receiver state machine) and K4 (sync search) on a config-4-shaped batch."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "ultrasonic-communication_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import numpy as np, torch
import usc, synth
from oracle import pyref as R

S = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
F = 381
N = 2048
dev = torch.device("cuda", 0)
h = usc.Handle()
st = torch.cuda.current_stream(); h.set_stream(st.cuda_stream)
base = synth.make_stream(b"Hello World!", snr_db=20.0, lead_in=40, guard=12, nframes=F)      # one 10 s stream
tb = torch.from_numpy(base.reshape(-1)).to(dev)
g = torch.Generator(device=dev); g.manual_seed(1)
offs = torch.randint(0, N, (S,), generator=g, device=dev)
pcm = torch.empty((S, F * N), dtype=torch.int32, device=dev)
idx = torch.arange(F * N, device=dev)
for s0 in range(0, S, 256):
    e = min(S, s0 + 256)
    sh = (idx[None, :] - offs[s0:e, None]) % (F * N)
    noise = (torch.randn((e - s0, F * N), generator=g, device=dev) * 2000).round().to(torch.int32) * 256
    pcm[s0:e] = tb[sh] + noise
uart = torch.zeros((S, 64), dtype=torch.uint8, device=dev)
res = torch.zeros((S, 8), dtype=torch.int32, device=dev)
def timeit(fn, reps=3):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    for _ in range(reps): fn()
    e1.record(st); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
ms = timeit(lambda: h.receiver_run(pcm, usc.PCM_I32, S, F, F * N, uart, 64, res))
u = uart.cpu().numpy(); r = res.cpu().numpy()
ok = sum(bytes(u[s, :13]) == b"Hello World!\n" for s in range(S))
print("K7 receiver_run: %d streams x %d frames: %.2f ms  %.1f Mframes/s (x4 FFT chains)  %.0f GB/s input  decoded %d/%d" % (S, F, ms, S * F / ms / 1e3, S * F * 8192 / ms / 1e6, ok, S))
# oracle check on a few streams
rx = R.RefReceiver()
for s in (0, 1, S - 1):
    want, stt = R.receiver_run(rx, pcm[s].cpu().numpy().reshape(F, N), cap=64)
    assert bytes(u[s, :min(r[s, 4], 64)]) == want[:64] and r[s, 2] == stt.lock_frame, s
mag = torch.empty((S, F, 4), dtype=torch.float32, device=dev); ii = torch.empty((S, F, 4), dtype=torch.int32, device=dev)
for K in (1, 4):
    ms = timeit(lambda: h.sync_search(pcm, usc.PCM_I32, S, F, F * N, K, mag, ii))
    print("K4 sync_search K=%d: %.2f ms  %.1f Mframes/s (x4 offsets)  %.0f GB/s input" % (K, ms, S * F / ms / 1e3, S * F * 8192 / ms / 1e6))
