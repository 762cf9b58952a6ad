"""Short driver for ncu captures of the synchroniser-side kernels: K7 k_receiver_run, K4 k_sync_search (K = 1, 4)
and K3 k_dsp2048c, on a config-4-shaped batch generated on the device.  `python tools/prof_rx.py rx|sync1|sync4|dspc [S]`."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "ultrasonic-communication_b200")):
    sys.path.insert(0, p)
import torch, usc

which = sys.argv[1] if len(sys.argv) > 1 else "rx"
S = int(sys.argv[2]) if len(sys.argv) > 2 else 2048
F, N, MB = 381, 2048, 12
dev = torch.device("cuda", 0)
if which == "dspc":
    B = 65536
    fifo = torch.randn((B, 3 * N), device=dev) * 1e4
    pos = torch.randint(0, 2 * N, (B,), dtype=torch.int32, device=dev)
    mean = torch.full((B,), 1e6, device=dev)
    hist = torch.empty((B, 12), dtype=torch.int32, device=dev)
    h = usc.Handle(usc.default_config(chirp_variant=usc.CHIRP_S))
    for _ in range(2):
        h.dsp(fifo, 3 * N, pos, mean, usc.UP, hist, B)
else:
    h = usc.Handle()
    pcm = torch.empty((S, F * N), dtype=torch.int32, device=dev)
    offs = torch.empty(S, dtype=torch.int32, device=dev)
    msgs = torch.empty((S, MB), dtype=torch.uint8, device=dev)
    h.synth_streams(4, 0, S, F, F * N, 40, MB, 12, 2.0e4, 2000.0, pcm, offs, msgs)
    if which == "rx":
        uart = torch.zeros((S, 64), dtype=torch.uint8, device=dev)
        res = torch.zeros((S, 8), dtype=torch.int32, device=dev)
        for _ in range(2):
            h.receiver_run(pcm, usc.PCM_I32, S, F, F * N, uart, 64, res)
    else:
        K = int(which[4:])
        mag = torch.empty((S, F, 4), dtype=torch.float32, device=dev)
        ii = torch.empty((S, F, 4), dtype=torch.int32, device=dev)
        for _ in range(2):
            h.sync_search(pcm, usc.PCM_I32, S, F, F * N, K, mag, ii)
torch.cuda.synchronize()
print("done", which)
