"""K7 (receiver state machine) and K4 (sync search) on a config-4-shaped batch generated on the device
(usc_synth_streams): S streams x 381 frames (10 s each), per-stream random start offset, random 12-byte
messages.  `python tools/quick_rx.py 32768` is the whole per-GPU shard of BASELINE config 4 (102 GB)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "ultrasonic-communication_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import numpy as np, torch
import usc
from oracle import pyref as R

S = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
F, N, MB = 381, 2048, 12
SEED, LEAD, GUARD, AMP, SIGMA = 4, 40, 12, 2.0e4, 2000.0
dev = torch.device("cuda", 0)
h = usc.Handle()
st = torch.cuda.current_stream(); h.set_stream(st.cuda_stream)
pcm = torch.empty((S, F * N), dtype=torch.int32, device=dev)
offs = torch.empty(S, dtype=torch.int32, device=dev)
msgs = torch.empty((S, MB), dtype=torch.uint8, device=dev)
def timeit(fn, reps=3):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    for _ in range(reps): fn()
    e1.record(st); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
ms = timeit(lambda: h.synth_streams(SEED, 0, S, F, F * N, LEAD, MB, GUARD, AMP, SIGMA, pcm, offs, msgs), reps=1)
print("synth_streams: %d streams x %d frames (%.1f GB): %.1f ms  %.0f GB/s written" % (S, F, S * F * N * 4 / 1e9, ms, S * F * N * 4 / ms / 1e6))
uart = torch.zeros((S, 64), dtype=torch.uint8, device=dev)
res = torch.zeros((S, 8), dtype=torch.int32, device=dev)
ms = timeit(lambda: h.receiver_run(pcm, usc.PCM_I32, S, F, F * N, uart, 64, res))
u = uart.cpu().numpy(); r = res.cpu().numpy(); m = msgs.cpu().numpy()
# the pattern (40 + 8 + 96 + 12 = 156 symbols) fits twice into 381 frames: the first message is at the start of uart
ok = sum(bytes(u[s, :MB + 1]) == bytes(m[s]) + b"\n" for s in range(S))
print("K7 receiver_run: %d streams x %d frames: %.2f ms  %.1f Mframes/s (x4 FFT chains)  %.0f GB/s input  decoded %d/%d" % (S, F, ms, S * F / ms / 1e3, S * F * 8192 / ms / 1e6, ok, S))
rx = R.RefReceiver()
for s in (0, 1, S - 1):                                   # oracle check on regenerated streams (nothing copied back)
    p1, _, _ = R.synth_streams(SEED, s, 1, F, LEAD, MB, GUARD, AMP, SIGMA)
    want, stt = R.receiver_run(rx, p1[0], cap=64)
    assert bytes(u[s, :min(r[s, 4], 64)]) == want[:64] and r[s, 2] == stt.lock_frame, s
print("K7 == oracle on streams 0, 1, %d (regenerated on the CPU)" % (S - 1))
if S <= 16384:
    mag = torch.empty((S, F, 4), dtype=torch.float32, device=dev); ii = torch.empty((S, F, 4), dtype=torch.int32, device=dev)
    for K in (1, 2, 4):
        ms = timeit(lambda: h.sync_search(pcm, usc.PCM_I32, S, F, F * N, K, mag, ii))
        print("K4 sync_search K=%d: %.2f ms  %.1f Mframes/s (x4 offsets)  %.0f GB/s input" % (K, ms, S * F / ms / 1e3, S * F * 8192 / ms / 1e6))
