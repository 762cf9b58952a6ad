"""Kernels that keep their tables in tensor memory, launched side by side on different streams (GPU).

Every fused kernel allocates TMEM columns per CTA (256 or 512 of an SM's 512): CTAs of different kernels that meet on
one SM must either share the columns or wait for each other, never dead-lock, and the results must equal those of the
same calls run one after the other."""
import os

import numpy as np
import pytest

import synth
import usc

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_tmem_kernels_on_concurrent_streams_match_sequential_runs():
    torch = pytest.importorskip("torch")
    dev = torch.device("cuda", 0)
    N, F = 2048, 4096
    taps = np.load(os.path.join(ROOT, "tests/golden/fir_taps.npz"))["taps"].astype(np.float32)[::-1].copy()
    h1, h5, h6 = usc.Handle(), usc.Handle(), usc.Handle(usc.default_config(n=8192))
    h5.iq_init(18000.0, 3000.0, taps, 32)
    pcm = torch.empty((F, N), dtype=torch.int32, device=dev)
    h1.synth_frames(7, 0, F, 2.0e4, 2.0e4, pcm)
    torch.cuda.synchronize()

    def outs():
        return ([torch.zeros(F, dtype=torch.float32, device=dev) for _ in range(2)] +
                [torch.zeros(F, dtype=torch.int32, device=dev) for _ in range(2)] + [torch.zeros(F, dtype=torch.uint8, device=dev)])

    def run(handle, kind, o):
        if kind == "k1":
            handle.demod_frames(pcm, usc.PCM_I32, F, o[0], o[2], o[1], o[3], o[4])
        elif kind == "k5":
            handle.iq_demod(pcm, usc.PCM_I32, F // 8, 8, 8 * N, o[0], o[2], o[1], o[3], o[4])
        else:                                               # the same memory seen as 8192-point frames
            handle.demod_frames(pcm, usc.PCM_I32, F // 4, o[0], o[2], o[1], o[3], o[4])

    jobs = [(h1, "k1"), (h5, "k5"), (h6, "k6")]
    seq = []
    for hd, kind in jobs:                                   # one after the other on the default stream
        o = outs()
        hd.set_stream(torch.cuda.current_stream().cuda_stream)
        run(hd, kind, o)
        torch.cuda.synchronize()
        seq.append([t.cpu().numpy() for t in o])
    streams = [torch.cuda.Stream(device=dev) for _ in jobs]
    par = [outs() for _ in jobs]
    torch.cuda.synchronize()
    for rep in range(6):                                    # interleaved launches: CTAs of the three kernels meet on the SMs
        for (hd, kind), st, o in zip(jobs, streams, par):
            hd.set_stream(st.cuda_stream)
            run(hd, kind, o)
    torch.cuda.synchronize()
    for s, p, (_, kind) in zip(seq, par, jobs):
        for a, b in zip(s, p):
            assert np.array_equal(a.view(np.uint8), b.cpu().numpy().view(np.uint8)), kind
    for hd, _ in jobs:
        hd.close()
