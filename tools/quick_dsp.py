"""Throughput of usc_dsp (gathered frames at per-stream sync positions): receiver variant R (k_dsp2048) and the
complex-FFT variant S of experiments/synchronization (k_dsp2048c)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "ultrasonic-communication_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import torch, usc
dev = torch.device("cuda", 0); st = torch.cuda.current_stream()
B, N = 65536, 2048
fifo = torch.randn((B, 3 * N), device=dev) * 1e4
pos = torch.randint(0, 2 * N, (B,), dtype=torch.int32, device=dev)
mean = torch.full((B,), 1e6, device=dev)
hist = torch.empty((B, 12), dtype=torch.int32, device=dev)
def timeit(fn, reps=10):
    for _ in range(3): fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record(st)
    for _ in range(reps): fn()
    e1.record(st); torch.cuda.synchronize(); return e0.elapsed_time(e1) / reps
for name, var in (("R (k_dsp2048)", usc.CHIRP_R), ("S (k_dsp2048c)", usc.CHIRP_S)):
    h = usc.Handle(usc.default_config(chirp_variant=var)); h.set_stream(st.cuda_stream)
    ms = timeit(lambda: h.dsp(fifo, 3 * N, pos, mean, usc.UP, hist, B))
    print("usc_dsp variant %-15s %d calls: %.3f ms  %.1f Mcalls/s  %.0f GB/s of frame bytes" % (name, B, ms, B / ms / 1e3, B * 8192 / ms / 1e6))
    h.close()
