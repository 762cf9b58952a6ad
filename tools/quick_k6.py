"""A/B loop for the long-frame kernels (K6): device-resident timing per frame length on 1 GiB of PCM.
usage: python tools/quick_k6.py [--sizes 8192,16384] [lib.so ...]   (each library in its own process via USC_LIB)"""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
args = sys.argv[1:]
sizes = "4096,8192,16384,32768,65536"
if args and args[0] == "--sizes":
    sizes = args[1]; args = args[2:]
if args and args[0] != "--child":
    for lib in args:
        env = dict(os.environ, USC_LIB=os.path.abspath(lib))
        r = subprocess.run([sys.executable, __file__, "--child", sizes], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        print("%-28s %s" % (os.path.basename(lib), r.stdout.strip().splitlines()[-1] if r.stdout.strip() else "no output"), flush=True)
    sys.exit(0)
if args and args[0] == "--child":
    sizes = args[1]
for p in (ROOT, os.path.join(ROOT, "ultrasonic-communication_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import torch
import usc
dev = torch.device("cuda", 0)
st = torch.cuda.current_stream()
def timeit(fn, reps=5):
    fn(); fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    for _ in range(reps): fn()
    e1.record(st); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
out = []
for n in [int(x) for x in sizes.split(",")]:
    hh = usc.Handle(usc.default_config(n=n)); hh.set_stream(st.cuda_stream)
    nf = (1 << 28) // n                       # 1 GiB of PCM
    x = torch.empty((nf, n), dtype=torch.int32, device=dev)
    hh.synth_frames(2, 0, nf, 2.0e4, 1.0e5, x)
    oo = [torch.empty(nf, dtype=torch.float32, device=dev) for _ in range(2)] + [torch.empty(nf, dtype=torch.int32, device=dev) for _ in range(2)]
    bb = torch.empty(nf, dtype=torch.uint8, device=dev)
    ms = min(timeit(lambda: hh.demod_frames(x, usc.PCM_I32, nf, oo[0], oo[2], oo[1], oo[3], bb)) for _ in range(2))
    out.append("%d: %.3f ms %.1f%%" % (n, ms, nf * (4 * n + 16) / ms / 1e6 / 65.52))
    hh.close()
print(" | ".join(out))
