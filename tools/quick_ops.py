"""Throughput of the batched CMSIS-shaped operators (the drop-in boundary) on the config-2 batch."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "ultrasonic-communication_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import torch, usc
dev = torch.device("cuda", 0); st = torch.cuda.current_stream()
h = usc.Handle(); h.set_stream(st.cuda_stream)
B, N = 155648, 2048
x = torch.randn((B, N), device=dev); y = torch.empty_like(x); w = torch.randn(N, device=dev)
m = torch.empty((B, N // 2), device=dev); v = torch.empty(B, device=dev); ix = torch.empty(B, dtype=torch.int32, device=dev)
xi = torch.randint(-2**30, 2**30, (B, N), dtype=torch.int32, device=dev)
def timeit(fn, reps=10):
    for _ in range(3): fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record(st)
    for _ in range(reps): fn()
    e1.record(st); torch.cuda.synchronize(); return e0.elapsed_time(e1) / reps
def rep(name, ms, bytes_):
    print("%-28s %.3f ms  %6.0f GB/s (%.0f%% of 6552)" % (name, ms, bytes_ / ms / 1e6, bytes_ / ms / 1e6 / 65.52))
G = B * N * 4
rep("i32_to_f32", timeit(lambda: h.i32_to_f32(xi, y, B * N)), 2 * G)
rep("arm_mult_f32 (bcast window)", timeit(lambda: h.arm_mult_f32(x, N, w, 0, y, N, N, B)), 2 * G)
rep("arm_cmplx_mult_cmplx_f32", timeit(lambda: h.arm_cmplx_mult_cmplx_f32(x, N, w, 0, y, N, N // 2, B)), 2 * G)
rep("arm_rfft_fast_f32 2048", timeit(lambda: h.arm_rfft_fast_f32(N, x, y, 0, B)), 2 * G)
rep("arm_rfft_fast_f32 2048 inv", timeit(lambda: h.arm_rfft_fast_f32(N, y, x, 1, B)), 2 * G)
rep("arm_cfft_f32 1024", timeit(lambda: h.arm_cfft_f32(1024, x, 0, B)), 2 * G)
rep("arm_cfft_f32 2048", timeit(lambda: h.arm_cfft_f32(2048, x, 0, B // 2)), 2 * G)
rep("arm_cmplx_mag_f32", timeit(lambda: h.arm_cmplx_mag_f32(y, N, m, N // 2, N // 2, B)), G + G // 2)
rep("arm_max_f32 (1024)", timeit(lambda: h.arm_max_f32(m, N // 2, N // 2, v, ix, B)), G // 2)
rep("arm_max_f32 (156 of 1024)", timeit(lambda: h.arm_max_f32(m, N // 2, 156, v, ix, B)), B * 156 * 4)
