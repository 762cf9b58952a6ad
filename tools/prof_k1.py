"""Short driver for ncu: a few launches of the fused kernels on the config-2 batch."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "ultrasonic-communication_b200"))
import torch  # noqa: E402
import usc  # noqa: E402
import bench  # noqa: E402

which = sys.argv[1] if len(sys.argv) > 1 else "demod"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
dev = torch.device("cuda", 0)
_h0 = usc.Handle()
pcm, bits = bench.make_device_frames(torch, _h0, bench.NFRAMES, dev, 0)
F = bench.NFRAMES
o = [torch.empty(F, dtype=torch.float32, device=dev) for _ in range(4)]
b = torch.empty(F, dtype=torch.uint8, device=dev)
if which == "demod":
    h = usc.Handle()
    for _ in range(reps):
        h.demod_frames(pcm, usc.PCM_I32, F, o[0], o[1], o[2], o[3], b)
elif which == "compress":
    h = usc.Handle(usc.default_config(fs=100000.0, f0=17000.0, f1=18000.0, chirp_variant=usc.CHIRP_T,
                                      window=usc.HANN_SYMMETRIC))
    for _ in range(reps):
        h.compress_chirp(pcm, usc.PCM_I32, F, False, None, o[0], o[1])
torch.cuda.synchronize()
print("done", which, reps)
if which == "single":
    h = usc.Handle()
    for _ in range(reps):
        h.demod_frames(pcm, usc.PCM_I32, F, o[0], o[1])
    torch.cuda.synchronize()
    print("done single", reps)
if which == "iq":
    import numpy as np
    taps = np.load(os.path.join(ROOT, "tests/golden/fir_taps.npz"))["taps"].astype(np.float32)[::-1].copy()
    h = usc.Handle()
    h.iq_init(18000.0, 3000.0, taps, 32)
    for _ in range(reps):
        h.iq_demod(pcm, usc.PCM_I32, 4096, 38, 38 * 2048, o[0], o[1], o[2], o[3], b)
    torch.cuda.synchronize()
    print("done iq", reps)
if which == "long32":
    hh = usc.Handle(usc.default_config(n=65536))
    nf = 2048
    x = torch.empty((nf, 65536), dtype=torch.int32, device=dev)
    hh.synth_frames(2, 0, nf, 2.0e4, 1.0e5, x)
    for _ in range(reps):
        hh.demod_frames(x, usc.PCM_I32, nf, o[0], o[1], o[2], o[3], b)
    torch.cuda.synchronize()
    print("done long32", reps)
if which == "legacy":
    h = usc.Handle()
    lv = torch.empty(F, dtype=torch.int8, device=dev); s16 = torch.empty(F, dtype=torch.int16, device=dev)
    for _ in range(reps):
        h.onoff_detect(pcm, usc.PCM_I32, 4096, 38, None, s16, lv)
    torch.cuda.synchronize()
    print("done legacy", reps)
if which in ("long8192", "long16384"):
    n = int(which[4:])
    hh = usc.Handle(usc.default_config(n=n))
    nf = (1 << 28) // n
    x = torch.empty((nf, n), dtype=torch.int32, device=dev)
    hh.synth_frames(2, 0, nf, 2.0e4, 1.0e5, x)
    for _ in range(reps):
        hh.demod_frames(x, usc.PCM_I32, nf, o[0], o[1], o[2], o[3], b)
    torch.cuda.synchronize()
    print("done", which, reps)
if which == "os":
    h = usc.Handle()
    S, Fr = 4096, 38
    mv = torch.empty(S * (Fr - 1), dtype=torch.float32, device=dev); mi = torch.empty(S * (Fr - 1), dtype=torch.int32, device=dev)
    for _ in range(reps):
        h.correlate_os(pcm, usc.PCM_I32, S, Fr, Fr * 2048, False, None, mv, mi)
    torch.cuda.synchronize()
    print("done os", reps)
if which.startswith("long:"):
    n = int(which.split(":")[1])
    hh = usc.Handle(usc.default_config(n=n))
    nf = (1 << 27) // n
    x = torch.empty((nf, n), dtype=torch.int32, device=dev)
    hh.synth_frames(2, 0, nf, 2.0e4, 1.0e5, x)
    for _ in range(reps):
        hh.demod_frames(x, usc.PCM_I32, nf, o[0], o[1], o[2], o[3], b)
    torch.cuda.synchronize()
    print("done", which, reps)
