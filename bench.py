#!/usr/bin/env python
"""bench.py — chirp symbols demodulated per second on B200 (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl native|reference]

Workload (BASELINE.json configs[1]): the receiver chain — int32 PCM -> de-chirp
(up AND down hypothesis) x Hann x 2048-pt RFFT x magnitude x arg-max over 156 bins x symbol decision —
over 4096 synthetic streams x 1 s @ 78 125 Hz = 155 648 frames (1.275 GB of PCM) per GPU.
A "step" is one pass of the hot path over that batch (one launch of the fused kernel K1).

  value     symbols/s with the PCM already resident in HBM (CUDA events on the launching stream)
  e2e       the same metric through the host-buffer C-ABI call usc_demod_frames_host(): pinned host
            PCM -> chunked H2D -> K1 -> D2H of the per-frame results, all inside the timed region
  roofline  algorithmic bytes (8208 B/symbol: 8192 in + 16 out) / average launch duration vs the
            measured HBM copy bandwidth (MEASURED_PEAKS.json)
  cpu_baseline  the CPU port of the same chain on this box's host cores on a bounded sample of the workload: the tuned
                form (oracle/ref_fast.c: one frame per SIMD lane, iterative plan, per-thread scratch, bit-identical to
                the plain restatement oracle/ref_dsp.c, whose rate is reported beside it)
  configs   one measured entry per further BASELINE.json config, same launch: "3" I/Q path (K5, 16384 streams),
            "4" synchroniser + receiver state machine (K7, and K4 with K = 1/2/4 added frames) on this rank's shard of
            262144/8 streams x 10 s, "5" 8192/16384/65536-point frames (K6) — each with value, ms, roofline, a sampled
            oracle check and an e2e figure through the host-buffer C-ABI call

`--impl reference` times the CPU implementation alone (the reference's C chain cannot be built for
the host: CMSIS-DSP is vendored only as an ARM archive; the tuned oracle port stands in).
Multi-GPU: one process per GPU (torchrun), streams sharded, no data-path collective, weak scaling.
"""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "ultrasonic-communication_b200"))

N = 2048
FS = 78125.0
F0, F1 = 16000.0, 19000.0
STREAMS = 4096
FRAMES_PER_STREAM = 38                      # 1 s @ 78 125 Hz = 38 whole frames
NFRAMES = STREAMS * FRAMES_PER_STREAM       # 155 648
ALGO_BYTES_PER_SYMBOL = N * 4 + 16
METRIC = "chirp symbols demodulated/sec"
UNIT = "symbols/s"
WORKLOAD = "receiver chain (Hanning + 2048-pt RFFT + chirp compression + peak), 4096 streams x 1 s"


SEED = 20261017
AMP = 2.0e4                                  # symbol amplitude before the x256 (|cos+sin| <= sqrt(2))
SNR_DB = -5.0
NOISE_SIGMA = AMP / (10.0 ** (SNR_DB / 20.0))    # chirp_orth has unit mean power


def make_device_frames(torch, h, nframes, device, first_frame):
    """Synthetic PCM generated ON THE DEVICE by the library's counter-based generator (usc_synth_frames:
    chirp_orth symbol chosen by a random bit + noise, int32 x256).  Frames are indexed globally, so
    rank r generates frames [r*nframes, (r+1)*nframes) of one reproducible dataset."""
    pcm = torch.empty((nframes, N), dtype=torch.int32, device=device)
    bits = torch.empty(nframes, dtype=torch.uint8, device=device)
    h.synth_frames(SEED, first_frame, nframes, AMP, NOISE_SIGMA, pcm, bits)
    h.sync()
    return pcm, bits


class ClockSampler:
    """SM clock / throttle reasons sampled through NVML DURING the timed region."""

    def __init__(self, index):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._t = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.dev = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.dev, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _run(self):
        nv = self.nv
        names = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20,
                 "hw_power_brake_slowdown": 0x80}
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.dev, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.dev)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.dev)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            time.sleep(0.002)

    def start(self):
        if self.nv:
            self._t = threading.Thread(target=self._run, daemon=True)
            self._t.start()

    def stop(self):
        self._stop.set()
        if self._t:
            self._t.join()
        s = sorted(self.samples)
        return {"sm_mhz": (s[len(s) // 2] if s else None), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(s)}


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def kernel_source_hash():
    """sha1 over the sources of the bench kernel: profiles/k1_traffic.json records it at capture time."""
    import hashlib
    hsh = hashlib.sha1()
    for f in ("k_demod.cu", "usc_warpfft.cuh", "usc_arith.cuh", "usc_tmem.cuh"):
        with open(os.path.join(ROOT, "ultrasonic-communication_b200", "csrc", f), "rb") as fh:
            hsh.update(fh.read())
    return hsh.hexdigest()[:16]


def profiled_traffic():
    """dram bytes per launch of K1 from the committed ncu --set full capture (ncu cannot run inside the bench);
    None when the kernel's sources changed since that capture (tools/update_traffic.py refreshes it)."""
    p = os.path.join(ROOT, "profiles", "k1_traffic.json")
    if os.path.exists(p):
        try:
            t = json.load(open(p))
            if t.get("kernel_source_sha1") == kernel_source_hash():
                return t
            return {"dram_bytes_per_launch": None, "stale": "kernel sources changed since the ncu capture"}
        except Exception:
            pass
    return None


def cpu_port(nframes_sample, threads, budget_s=10.0, fast=True):
    """The CPU port (test infrastructure) timed as the reported baseline: passes over a bounded sample of the same
    dataset until the budget is spent.  fast: the tuned form (ref_fast.c); else the plain restatement (ref_dsp.c)."""
    from oracle import pyref
    rx = pyref.RefReceiver()
    pcm, _ = pyref.synth_frames(SEED, 0, nframes_sample, AMP, NOISE_SIGMA)      # the same dataset, first frames
    run = rx.demod_frames_fast if fast else rx.demod_frames
    run(pcm[:512], nthreads=threads)                # warm the threads
    t_start = time.perf_counter()
    passes = 0
    while True:
        run(pcm, nthreads=threads)
        passes += 1
        elapsed = time.perf_counter() - t_start
        if elapsed >= budget_s or passes >= 4096:
            break
    return nframes_sample * passes / elapsed, elapsed, passes


def cpu_numpy(nframes_sample, threads, budget_s=4.0):
    """The chain as the reference's Python simulations compose it (simulation/dsp.py:83-87): numpy elementwise +
    scipy FFT + abs + arg-max per hypothesis, float32, all cores.  A second reported CPU number."""
    import scipy.fft
    from oracle import pyref
    rx = pyref.RefReceiver()
    up, down, hann = rx.table("up_chirp"), rx.table("down_chirp"), rx.table("hann")
    bw2 = rx.bandwidth2
    pcm, _ = pyref.synth_frames(SEED, 0, nframes_sample, AMP, NOISE_SIGMA)

    def once():
        x = pcm.astype(np.float32)
        out = []
        for c in (up, down):
            m = np.abs(scipy.fft.rfft(x * c * hann, axis=1, workers=threads)[:, :bw2])
            out.append((m.max(axis=1), m.argmax(axis=1)))
        return out

    once()
    t0 = time.perf_counter()
    passes = 0
    while True:
        once()
        passes += 1
        dt = time.perf_counter() - t0
        if dt >= budget_s:
            break
    return nframes_sample * passes / dt, dt, passes


def cpu_simd_note():
    from oracle import pyref
    w = pyref.lib().ref_fast_simd_width()
    return {16: "AVX-512, 16 frames per vector", 8: "AVX2, 8 frames per vector"}.get(w, "scalar")


def run_reference(args, rank):
    """--impl reference: the CPU implementation alone, all host threads, bounded sample per step."""
    if rank != 0:
        return
    threads = len(os.sched_getaffinity(0))
    sample = 32768
    from oracle import pyref
    rx = pyref.RefReceiver()
    pcm, _ = pyref.synth_frames(SEED, 0, sample, AMP, NOISE_SIGMA)
    for _ in range(max(args.warmup, 1)):
        rx.demod_frames_fast(pcm[:4096], nthreads=threads)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        rx.demod_frames_fast(pcm, nthreads=threads)
    dt = time.perf_counter() - t0
    val = sample * args.steps / dt
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "frames_per_step": sample, "n": N, "hypotheses": 2,
                   "note": "tuned CPU port of the receiver chain (reference C chain not buildable on host: "
                           "CMSIS-DSP vendored only as an ARM-Thumb archive); bounded sample of the workload per step"},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": "%d frames x %d steps, OpenMP over blocks of frames, %s" % (sample, args.steps, cpu_simd_note())},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ---- the further BASELINE.json configs (3, 4, 5): measured in the same launch, reported under "configs" ----------------
def _dev_ms(torch, stream, fn, reps, warm):
    for _ in range(warm):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record(stream)
    for _ in range(reps):
        fn()
    e1.record(stream)
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def _wall_s(torch, barrier, fn, reps):
    fn()
    barrier()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    barrier()
    return (time.perf_counter() - t0) / reps


IQ_CARRIER, IQ_BW = 18000.0, 3000.0
C3_STREAMS, C3_E2E_STREAMS = 16384, 2048
C4_STREAMS_TOTAL, C4_FRAMES, C4_MSG, C4_LEAD, C4_GUARD, C4_SIGMA = 262144, 381, 12, 40, 12, 2000.0
C4_E2E_STREAMS = 384                     # 384 x 381 frames = 1.2 GB: fits the pinned arena of config 2's e2e leg
C5_LENGTHS = ((8192, -10.0), (16384, -15.0), (65536, -20.0))   # SNRs at which the longer chirps still decode (accuracy is reported)


def config3(torch, usc, pyref, dev, stream, rank, barrier, arena):
    """BASELINE config 3: the I/Q baseband path (K5) over 16384 streams x 38 frames per GPU.  Input: the I/Q transmitter's
    symbols (usc_synth_iq_frames: A cos(2 pi (fc - fb(t)) t), simulation/IQ_modulation.ipynb cell 4) + noise, SNR 0 dB."""
    taps = np.load(os.path.join(ROOT, "tests", "golden", "fir_taps.npz"))["taps"].astype(np.float32)[::-1].copy()
    h = usc.Handle(device=dev.index)
    h.set_stream(stream.cuda_stream)
    h.iq_init(IQ_CARRIER, IQ_BW, taps, 32)
    S, F = C3_STREAMS, FRAMES_PER_STREAM
    nfr = S * F
    sigma = AMP / np.sqrt(2.0)                               # a cosine of amplitude A has power A^2/2: SNR 0 dB
    pcm = torch.empty((nfr, N), dtype=torch.int32, device=dev)
    bits = torch.empty(nfr, dtype=torch.uint8, device=dev)
    first = rank * nfr
    h.synth_iq_frames(SEED + 3, first, nfr, IQ_CARRIER, IQ_BW, -1, 0.0, AMP, sigma, pcm, bits)
    o = [torch.empty(nfr, dtype=torch.float32, device=dev) for _ in range(2)] + \
        [torch.empty(nfr, dtype=torch.int32, device=dev) for _ in range(2)]
    b = torch.empty(nfr, dtype=torch.uint8, device=dev)
    ms = _dev_ms(torch, stream, lambda: h.iq_demod(pcm, usc.PCM_I32, S, F, F * N, o[0], o[2], o[1], o[3], b), reps=5, warm=2)
    acc = float((b == bits).float().mean().item())
    # sampled oracle check: two streams of this rank's shard regenerated on the CPU
    q = pyref.RefIq(taps)
    ok = True
    for sidx in (0, S - 1):
        p1, _ = pyref.synth_iq_frames(SEED + 3, first + sidx * F, F, IQ_CARRIER, IQ_BW, -1, 0.0, AMP, sigma)
        want = q.demod(p1)
        sl = slice(sidx * F, (sidx + 1) * F)
        ok &= bool(np.array_equal(o[0][sl].cpu().numpy().view(np.uint32), want[0].view(np.uint32)) and
                   np.array_equal(o[2][sl].cpu().numpy().astype(np.uint32), want[1]) and
                   np.array_equal(o[1][sl].cpu().numpy().view(np.uint32), want[2].view(np.uint32)))
    # e2e through usc_iq_demod_host on a bounded sample of the shard
    Se = C3_E2E_STREAMS
    ne = Se * F
    hp = arena[:ne * N].view(ne, N)
    hp.copy_(pcm[:ne])
    ho = [torch.empty(ne, dtype=torch.float32).pin_memory() for _ in range(2)] + \
         [torch.empty(ne, dtype=torch.int32).pin_memory() for _ in range(2)]
    hb = torch.empty(ne, dtype=torch.uint8).pin_memory()
    e2e_s = _wall_s(torch, barrier, lambda: h.iq_demod_hostbuf(hp, usc.PCM_I32, Se, F, F * N, ho[0], ho[2], ho[1], ho[3], hb), 2)
    e2e_ok = bool(torch.equal(hb, b[:ne].cpu()) and torch.equal(ho[2], o[2][:ne].cpu()))
    h.close()
    return {"times": {"c3_ms": ms, "c3_e2e_s": e2e_s}, "ok": ok and e2e_ok,
            "info": {"frames": nfr, "e2e_frames": ne, "accuracy": acc}}


def config4(torch, usc, pyref, dev, stream, rank, world, barrier, arena):
    """BASELINE config 4: 262144 streams x 10 s sharded over 8 GPUs -> 32768 streams x 381 frames per GPU (102 GB), one wave.
    K7 = the receiver's whole main loop (sliding-correlation search, lock, decode); K4 = the search grid alone with
    synchronous addition of K = 1, 2, 4 frames."""
    h = usc.Handle(device=dev.index)
    h.set_stream(stream.cuda_stream)
    S, F, MB = C4_STREAMS_TOTAL // 8, C4_FRAMES, C4_MSG
    first = rank * S
    pcm = torch.empty((S, F * N), dtype=torch.int32, device=dev)
    msgs = torch.empty((S, MB), dtype=torch.uint8, device=dev)
    h.synth_streams(SEED + 4, first, S, F, F * N, C4_LEAD, MB, C4_GUARD, AMP, C4_SIGMA, pcm, None, msgs)
    uart = torch.zeros((S, 64), dtype=torch.uint8, device=dev)
    res = torch.zeros((S, 8), dtype=torch.int32, device=dev)
    times = {"c4_k7_ms": _dev_ms(torch, stream, lambda: h.receiver_run(pcm, usc.PCM_I32, S, F, F * N, uart, 64, res), reps=2, warm=1)}
    u, m, r = uart.cpu().numpy(), msgs.cpu().numpy(), res.cpu().numpy()
    decoded = int(sum(bytes(u[s, :MB + 1]) == bytes(m[s]) + b"\n" for s in range(S)))
    rx = pyref.RefReceiver()
    ok = True
    for sidx in (0, S - 1):                                  # oracle on regenerated streams: bytes and lock frame
        p1, _, _ = pyref.synth_streams(SEED + 4, first + sidx, 1, F, C4_LEAD, MB, C4_GUARD, AMP, C4_SIGMA)
        want, stt = pyref.receiver_run(rx, p1[0], cap=64)
        ok &= bool(bytes(u[sidx, :min(int(r[sidx, 4]), 64)]) == want[:64] and int(r[sidx, 2]) == stt.lock_frame)
    mag = torch.empty((S, F, 4), dtype=torch.float32, device=dev)
    idx = torch.empty((S, F, 4), dtype=torch.int32, device=dev)
    p1, _, _ = pyref.synth_streams(SEED + 4, first + 1, 1, F, C4_LEAD, MB, C4_GUARD, AMP, C4_SIGMA)
    for K in (1, 2, 4):
        times["c4_k4_K%d_ms" % K] = _dev_ms(torch, stream, lambda: h.sync_search(pcm, usc.PCM_I32, S, F, F * N, K, mag, idx), reps=2, warm=1)
        wm, wi = pyref.sync_search(rx, p1[0], K)
        ok &= bool(np.array_equal(idx[1].cpu().numpy().astype(np.uint32), wi) and
                   np.array_equal(mag[1].cpu().numpy().view(np.uint32), wm.view(np.uint32)))
    del mag, idx
    # K8: the overlap-save synchroniser (every lag, each sample read once) on the same shard
    omv = torch.empty(S * (F - 1), dtype=torch.float32, device=dev)
    omi = torch.empty(S * (F - 1), dtype=torch.int32, device=dev)
    times["c4_os_ms"] = _dev_ms(torch, stream, lambda: h.correlate_os(pcm, usc.PCM_I32, S, F, F * N, False, None, omv, omi), reps=2, warm=1)
    tmpl = pyref.arm_mult_f32(h.table("down"), h.table("hann"))
    _, wv, wi = pyref.correlate_os(tmpl, p1[0].reshape(F, N))
    sl = slice(1 * (F - 1), 2 * (F - 1))
    ok &= bool(np.array_equal(omi[sl].cpu().numpy().astype(np.uint32), wi) and np.array_equal(omv[sl].cpu().numpy().view(np.uint32), wv.view(np.uint32)))
    del omv, omi
    # e2e through usc_receiver_run_host on a bounded sample of the shard
    Se = C4_E2E_STREAMS
    hp = arena[:Se * F * N].view(Se, F * N)
    hp.copy_(pcm[:Se])
    hu = torch.zeros((Se, 64), dtype=torch.uint8).pin_memory()
    hr = torch.zeros((Se, 8), dtype=torch.int32).pin_memory()
    h.host_workspace(32 * Se)                                # 32 frames of every stream per chunk
    times["c4_e2e_s"] = _wall_s(torch, barrier, lambda: h.receiver_run_hostbuf(hp, usc.PCM_I32, Se, F, F * N, hu, 64, hr), 1)
    e2e_ok = bool(np.array_equal(hu.numpy(), u[:Se]) and np.array_equal(hr.numpy(), r[:Se]))
    h.close()
    del pcm
    torch.cuda.empty_cache()
    return {"times": times, "ok": ok and e2e_ok, "info": {"streams": S, "frames_per_stream": F, "decoded": decoded, "e2e_streams": Se}}


def config5(torch, usc, pyref, dev, stream, rank, barrier, arena):
    """BASELINE config 5: 8192 / 16384 / 65536-point chirp frames at low SNR, total samples per GPU = config 2's."""
    times, info, ok = {}, {}, True
    for n, snr in C5_LENGTHS:
        h = usc.Handle(usc.default_config(n=n, sweep_T=n / FS), device=dev.index)   # the chirp fills the frame, as the generator's symbols do
        h.set_stream(stream.cuda_stream)
        nf = (NFRAMES * N) // n
        sigma = AMP * 10.0 ** (-snr / 20.0)
        x = torch.empty((nf, n), dtype=torch.int32, device=dev)
        bits = torch.empty(nf, dtype=torch.uint8, device=dev)
        first = rank * nf
        h.synth_frames(SEED + 5, first, nf, AMP, sigma, x, bits)
        o = [torch.empty(nf, dtype=torch.float32, device=dev) for _ in range(2)] + \
            [torch.empty(nf, dtype=torch.int32, device=dev) for _ in range(2)]
        b = torch.empty(nf, dtype=torch.uint8, device=dev)
        times["c5_%d_ms" % n] = _dev_ms(torch, stream, lambda: h.demod_frames(x, usc.PCM_I32, nf, o[0], o[2], o[1], o[3], b), reps=5, warm=2)
        p1, _ = pyref.synth_frames(SEED + 5, first + nf - 1, 1, AMP, sigma, n=n)
        want = pyref.RefReceiver(n=n, sweep_T=n / FS).demod_frames(p1, nthreads=1)
        ok &= bool(o[0][nf - 1].item() == want[0][0] and o[2][nf - 1].item() == want[1][0] and
                   o[1][nf - 1].item() == want[2][0] and o[3][nf - 1].item() == want[3][0])
        hp = arena[:nf * n].view(nf, n)
        hp.copy_(x)
        ho = [torch.empty(nf, dtype=torch.float32).pin_memory() for _ in range(2)] + \
             [torch.empty(nf, dtype=torch.int32).pin_memory() for _ in range(2)]
        hb = torch.empty(nf, dtype=torch.uint8).pin_memory()
        times["c5_%d_e2e_s" % n] = _wall_s(torch, barrier, lambda: h.demod_frames_hostbuf(hp, usc.PCM_I32, nf, ho[0], ho[2], ho[1], ho[3], hb), 2)
        ok &= bool(torch.equal(hb, b.cpu()) and torch.equal(ho[2], o[2].cpu()))
        info[str(n)] = {"frames": nf, "snr_db": snr, "accuracy": float((b == bits).float().mean().item())}
        h.close()
        del x
    return {"times": times, "ok": ok, "info": info}


def configs_report(t, ok, infos, world, peak):
    """the "configs" object of the JSON line from the rank-maximal timings"""
    def roof(bytes_per_gpu, ms):
        a = bytes_per_gpu / (ms * 1e-3) / 1e9
        return {"bound": "hbm", "achieved": a, "peak": peak, "unit": "GB/s", "frac": a / peak, "algorithmic_bytes": bytes_per_gpu}
    out = {}
    i3 = infos["3"]
    out["3"] = {
        "workload": "I/Q baseband path: carrier mix + 2 x 27-tap FIR + decimate by 2 + 1024-pt complex FFT, both hypotheses, %d streams x %d frames per GPU"
                    % (C3_STREAMS, FRAMES_PER_STREAM),
        "kernel": "k_iq_fused<int>", "value": i3["frames"] * world / (t["c3_ms"] * 1e-3), "unit": UNIT, "ms": t["c3_ms"],
        "roofline": roof(i3["frames"] * (N * 4 + 24), t["c3_ms"]),
        "e2e": {"value": i3["e2e_frames"] * world / t["c3_e2e_s"], "unit": UNIT, "call": "usc_iq_demod_host",
                "h2d_bytes_per_step": i3["e2e_frames"] * N * 4, "d2h_bytes_per_step": i3["e2e_frames"] * 17,
                "sample": "%d streams of the shard" % C3_E2E_STREAMS},
        "input": "usc_synth_iq_frames: A cos(2 pi (fc - fb(t)) t) + noise, SNR 0 dB (simulation/IQ_modulation.ipynb cell 4)",
        "symbol_accuracy_vs_tx_bits": i3["accuracy"], "oracle_check": ok["3"]}
    i4 = infos["4"]
    fr4 = i4["streams"] * i4["frames_per_stream"]
    c4 = {
        "workload": "time-frame synchronisation + receiver state machine: %d streams x 10 s sharded over 8 GPUs -> %d streams x %d frames (%.0f GB) per GPU, one wave"
                    % (C4_STREAMS_TOTAL, i4["streams"], i4["frames_per_stream"], fr4 * N * 4 / 1e9),
        "kernel": "k_receiver_run<int>", "value": fr4 * world / (t["c4_k7_ms"] * 1e-3), "unit": UNIT, "ms": t["c4_k7_ms"],
        "roofline": roof(fr4 * N * 4 + i4["streams"] * 96, t["c4_k7_ms"]),
        "messages_decoded_rank0": i4["decoded"], "streams_per_gpu": i4["streams"],
        "sync_search": {}, "oracle_check": ok["4"],
        "e2e": {"value": i4["e2e_streams"] * i4["frames_per_stream"] * world / t["c4_e2e_s"], "unit": UNIT,
                "call": "usc_receiver_run_host", "h2d_bytes_per_step": i4["e2e_streams"] * i4["frames_per_stream"] * N * 4,
                "d2h_bytes_per_step": i4["e2e_streams"] * 96, "sample": "%d streams of the shard" % C4_E2E_STREAMS}}
    for K in (1, 2, 4):
        ms = t["c4_k4_K%d_ms" % K]
        c4["sync_search"]["K%d" % K] = {"kernel": "k_sync_search", "frames_added": K, "value": fr4 * world / (ms * 1e-3), "unit": "frames/s",
                                        "ms": ms, "roofline": roof(fr4 * (N * 4 + 32), ms)}
    ms = t["c4_os_ms"]
    c4["overlap_save"] = {"kernel": "k_correlate_os<int, false>", "call": "usc_correlate_os", "what": "linear matched filter at every lag (2n windows, hop n, 4096-point real transforms), peak per block",
                          "value": i4["streams"] * (i4["frames_per_stream"] - 1) * world / (ms * 1e-3), "unit": "blocks/s", "ms": ms,
                          "roofline": roof(fr4 * N * 4 + i4["streams"] * (i4["frames_per_stream"] - 1) * 8, ms)}
    out["4"] = c4
    c5 = {"workload": "long chirp frames at low SNR, %d samples per GPU per pass" % (NFRAMES * N), "by_n": {}, "oracle_check": ok["5"]}
    for n, _ in C5_LENGTHS:
        i5 = infos["5"][str(n)]
        ms = t["c5_%d_ms" % n]
        c5["by_n"][str(n)] = {
            "kernel": "k_demod_long" if n <= 16384 else "k_demod_cluster", "frames": i5["frames"], "snr_db": i5["snr_db"],
            "value": i5["frames"] * world / (ms * 1e-3), "unit": UNIT, "msamples_per_s": i5["frames"] * world * n / (ms * 1e-3) / 1e6,
            "ms": ms, "roofline": roof(i5["frames"] * (4 * n + 16), ms), "symbol_accuracy_vs_tx_bits": i5["accuracy"],
            "e2e": {"value": i5["frames"] * world / t["c5_%d_e2e_s" % n], "unit": UNIT, "call": "usc_demod_frames_host",
                    "h2d_bytes_per_step": i5["frames"] * n * 4, "d2h_bytes_per_step": i5["frames"] * 17}}
    out["5"] = c5
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-configs", action="store_true", help="skip the BASELINE config 3/4/5 measurements")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference(args, rank)
        return

    import torch
    import torch.distributed as dist
    import usc

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    h = usc.Handle(device=local_rank)
    stream = torch.cuda.current_stream()
    h.set_stream(stream.cuda_stream)

    # this rank's shard of the streams (weak scaling: every GPU gets the config-2 batch)
    pcm, bits = make_device_frames(torch, h, NFRAMES, dev, first_frame=rank * NFRAMES)
    mag_up = torch.empty(NFRAMES, dtype=torch.float32, device=dev)
    mag_dn = torch.empty(NFRAMES, dtype=torch.float32, device=dev)
    idx_up = torch.empty(NFRAMES, dtype=torch.int32, device=dev)
    idx_dn = torch.empty(NFRAMES, dtype=torch.int32, device=dev)
    bit = torch.empty(NFRAMES, dtype=torch.uint8, device=dev)

    def step():
        h.demod_frames(pcm, usc.PCM_I32, NFRAMES, mag_up, idx_up, mag_dn, idx_dn, bit)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    l0 = h.launch_count
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(args.steps):
        step()
    e1.record(stream)
    barrier()
    clocks = sampler.stop()
    ms_total = e0.elapsed_time(e1)
    launches = h.launch_count - l0

    # secondary measurement (not the headline): ONE hypothesis per frame = the reference's dsp() call
    # (receiver/Src/main.c:183-215), the fused window+FFT+compression+peak kernel in pair mode
    for _ in range(3):
        h.demod_frames(pcm, usc.PCM_I32, NFRAMES, mag_up, idx_up)
    s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    s0.record(stream)
    for _ in range(args.steps):
        h.demod_frames(pcm, usc.PCM_I32, NFRAMES, mag_up, idx_up)
    s1.record(stream)
    torch.cuda.synchronize()
    ms_single = s0.elapsed_time(s1) / args.steps
    # reported comparison only (never on the product path): the same chain on library kernels — torch
    # element-wise ops + cuFFT (torch.fft.rfft) + abs + arg-max, >= 6 HBM passes per hypothesis
    cufft_ms = None
    if rank == 0:
        try:
            t_up = torch.from_numpy(h.table("up")).to(dev)
            t_dn = torch.from_numpy(h.table("down")).to(dev)
            t_w = torch.from_numpy(h.table("hann")).to(dev)
            sub = 16384                                      # frames per library batch (134 MB of PCM)

            def lib_chain():
                for f0 in range(0, NFRAMES, sub):
                    x = pcm[f0:f0 + sub].to(torch.float32)
                    for tab in (t_up, t_dn):
                        m = torch.fft.rfft(x * tab * t_w, dim=1)[:, :156].abs()
                        m.max(dim=1)
            lib_chain()
            torch.cuda.synchronize()
            c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            c0.record(stream)
            for _ in range(3):
                lib_chain()
            c1.record(stream)
            torch.cuda.synchronize()
            cufft_ms = c0.elapsed_time(c1) / 3
        except Exception as exc:                             # comparison is optional
            cufft_ms = None
            sys.stderr.write("cuFFT comparison skipped: %r\n" % (exc,))
    for _ in range(2):                                   # leave the dual-hypothesis results in the buffers
        step()
    torch.cuda.synchronize()
    accuracy = float((bit == bits).float().mean().item())

    # end-to-end through the host-buffer C-ABI call: pinned host PCM -> H2D -> K1 -> D2H results.
    # The host buffers are first touched from CPUs next to this rank's GPU (NVML's ideal affinity), so that with
    # several ranks each PCIe link is fed from its own NUMA node; the previous affinity comes back afterwards.
    old_affinity, numa_note = os.sched_getaffinity(0), "unchanged"
    try:
        import pynvml
        pynvml.nvmlInit()
        pynvml.nvmlDeviceSetCpuAffinity(pynvml.nvmlDeviceGetHandleByIndex(local_rank))
        numa_note = "host buffers touched from %d CPUs local to GPU %d" % (len(os.sched_getaffinity(0)), local_rank)
    except Exception as exc:                             # cpuset of the container may exclude them
        numa_note = "NVML affinity not applied (%s)" % type(exc).__name__
    host_pcm = torch.empty((NFRAMES, N), dtype=torch.int32).pin_memory()
    host_pcm.copy_(pcm)
    h_mu = torch.empty(NFRAMES, dtype=torch.float32).pin_memory()
    h_md = torch.empty(NFRAMES, dtype=torch.float32).pin_memory()
    h_iu = torch.empty(NFRAMES, dtype=torch.int32).pin_memory()
    h_id = torch.empty(NFRAMES, dtype=torch.int32).pin_memory()
    h_bit = torch.empty(NFRAMES, dtype=torch.uint8).pin_memory()
    h.host_workspace(4096)

    def e2e_step():
        h.demod_frames_hostbuf(host_pcm, usc.PCM_I32, NFRAMES, h_mu, h_iu, h_md, h_id, h_bit)   # blocking

    e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.e2e_steps):
        e2e_step()
    barrier()
    e2e_s = time.perf_counter() - t0
    e2e_ok = bool(torch.equal(h_bit, bit.cpu()) and torch.equal(h_iu, idx_up.cpu()))
    # the ceiling of that path: bare pinned host -> device copies of the same buffer, all ranks at once
    d_tmp = torch.empty_like(pcm)
    d_tmp.copy_(host_pcm, non_blocking=True)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.e2e_steps):
        d_tmp.copy_(host_pcm, non_blocking=True)
    barrier()
    bare_s = time.perf_counter() - t0
    del d_tmp

    # the further BASELINE configs (3, 4, 5), measured on this rank's shard in the same launch
    cfg_times, cfg_ok, cfg_infos = {}, {}, {}
    if not args.no_configs:
        from oracle import pyref                              # sampled oracle checks only (the checker, never the timed path)
        arena = host_pcm.view(-1)
        for key, fn in (("3", lambda: config3(torch, usc, pyref, dev, stream, rank, barrier, arena)),
                        ("5", lambda: config5(torch, usc, pyref, dev, stream, rank, barrier, arena)),
                        ("4", lambda: config4(torch, usc, pyref, dev, stream, rank, world, barrier, arena))):
            r = fn()
            cfg_times.update(r["times"])
            cfg_ok[key], cfg_infos[key] = r["ok"], r["info"]
    os.sched_setaffinity(0, old_affinity)
    del host_pcm

    keys = sorted(cfg_times)
    times = torch.tensor([ms_total, e2e_s, ms_single, bare_s] + [cfg_times[k] for k in keys], dtype=torch.float64, device=dev)
    oks = torch.tensor([int(e2e_ok)] + [int(cfg_ok[k]) for k in sorted(cfg_ok)], dtype=torch.int32, device=dev)
    if world > 1:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)          # timing only; no data-path collective
        dist.all_reduce(oks, op=dist.ReduceOp.MIN)
    ms_total, e2e_s, ms_single, bare_s = (float(v) for v in times[:4].tolist())
    cfg_times = dict(zip(keys, (float(v) for v in times[4:].tolist())))
    e2e_ok = bool(oks[0].item())
    cfg_ok = dict(zip(sorted(cfg_ok), (bool(v) for v in oks[1:].tolist())))
    # the ranks are done with each other: the CPU legs below run on rank 0 alone with no rank spinning in a barrier
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank != 0:
        return

    if rank == 0:
        ms_per_step = ms_total / args.steps
        value = NFRAMES * world * args.steps / (ms_total * 1e-3)
        peak, peak_src = measured_peak()
        achieved = ALGO_BYTES_PER_SYMBOL * NFRAMES / (ms_per_step * 1e-3) / 1e9          # per GPU
        traffic = profiled_traffic()
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "msamples_per_s": value * N / 1e6,
            "config": {"workload": WORKLOAD, "streams_per_gpu": STREAMS, "frames_per_stream": FRAMES_PER_STREAM,
                       "frames_per_step_per_gpu": NFRAMES, "n": N, "fs": FS, "pcm": "int32 (x256 DFSDM words)",
                       "hypotheses": 2, "snr_db": SNR_DB, "generator": "usc_synth_frames (Philox-4x32-10, CPU twin in oracle)", "parallelism": "streams sharded, no collective",
                       "l2": "input 1.275 GB per step >> 126 MB L2 (no flush needed)",
                       "symbol_accuracy_vs_tx_bits": accuracy},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "frac_of_nominal_8000": achieved / 8000.0,
                         "traffic": (traffic or {}).get("dram_bytes_per_launch"), "peak_source": peak_src,
                         "kernel": "k_demod2048<int,5,12>", "algorithmic_bytes_per_launch": ALGO_BYTES_PER_SYMBOL * NFRAMES,
                         "note": "tables served from tensor memory (tcgen05.ld), 12 warps per SM; bound by the fp32 pipe (77 % busy, 1034 packed FP instructions per frame), L1/shared path 55 %, DESIGN.md 4.1; frac is vs HBM"},
            # secondary roofline: algorithmic fp32 work of two real-FFT chains per symbol, 2.5 N log2 N + 8 N flops
            # each, against the fp32 FMA peak of the measured SM clock (148 SMs x 128 lanes x 2 flops)
            "roofline_fp32": {"bound": "fp32", "achieved": 2 * (2.5 * N * 11 + 8 * N) * NFRAMES / (ms_per_step * 1e-3) / 1e12,
                              "peak": 148 * 128 * 2 * (clocks.get("sm_mhz") or 1965) * 1e6 / 1e12, "unit": "TFLOP/s",
                              "frac": 2 * (2.5 * N * 11 + 8 * N) * NFRAMES / (ms_per_step * 1e-3) / (148 * 128 * 2 * (clocks.get("sm_mhz") or 1965) * 1e6),
                              "note": "nominal flop count of the textbook radix-2 chain; the kernel executes fewer (pruned last pass, trivial twiddles)"},
            "e2e": {"value": NFRAMES * world * args.e2e_steps / e2e_s, "unit": UNIT,
                    "h2d_bytes_per_step": NFRAMES * N * 4, "d2h_bytes_per_step": NFRAMES * 17,
                    "steps": args.e2e_steps, "results_match_device_path": e2e_ok, "host_numa": numa_note,
                    "h2d_gbs_achieved": NFRAMES * world * args.e2e_steps * N * 4 / e2e_s / 1e9,
                    "bare_pinned_h2d_gbs": NFRAMES * world * args.e2e_steps * N * 4 / bare_s / 1e9,
                    "fraction_of_bare_copy": bare_s / e2e_s,
                    "note": "bare_pinned_h2d_gbs = the same pinned buffers copied host -> device by all ranks at once with no "
                            "kernel: the ceiling of the host-buffer path on this box"},
            "roofline_single_hypothesis": {
                "bound": "hbm", "achieved": (N * 4 + 8) * NFRAMES / (ms_single * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                "frac": (N * 4 + 8) * NFRAMES / (ms_single * 1e-3) / 1e9 / peak,
                "frac_of_nominal_8000": (N * 4 + 8) * NFRAMES / (ms_single * 1e-3) / 1e9 / 8000.0, "kernel": "k_demod2048_pair<int,5>",
                "ms_per_launch": ms_single, "frames_per_s": NFRAMES / (ms_single * 1e-3),
                "note": "dsp() for one hypothesis (window + 2048-pt RFFT + compression + peak), two frames per warp"},
            "cufft_comparison": (None if cufft_ms is None else {
                "value": NFRAMES / (cufft_ms * 1e-3), "unit": UNIT, "ms_per_step": cufft_ms,
                "what": "torch elementwise + torch.fft.rfft (cuFFT) + abs + max per hypothesis on the same device data; "
                        "reported comparison, not the product path"}),
            "gpu_launches": launches,
            "clocks": clocks,
        }
        if not args.no_cpu_baseline:
            threads = len(os.sched_getaffinity(0))
            v, dt, passes = cpu_port(32768, threads, budget_s=10.0, fast=True)
            line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": threads, "kind": "port",
                                    "sample": "%d passes over 32768 frames of the same workload, OpenMP over blocks of frames, %s, %.1f s of wall time"
                                              % (passes, cpu_simd_note(), dt),
                                    "form": "oracle/ref_fast.c: one frame per SIMD lane, iterative plan, per-thread scratch; bit-identical to oracle/ref_dsp.c"}
            pv, pdt, ppasses = cpu_port(8192, threads, budget_s=4.0, fast=False)
            line["cpu_baseline"]["plain_restatement"] = {
                "value": pv, "unit": UNIT, "cores": threads,
                "sample": "%d passes over 8192 frames, oracle/ref_dsp.c (the readable scalar restatement), %.1f s" % (ppasses, pdt)}
            nv, ndt, npasses = cpu_numpy(8192, threads)
            line["cpu_baseline"]["numpy_chain"] = {
                "value": nv, "unit": UNIT, "cores": threads,
                "sample": "%d passes over 8192 frames, scipy.fft.rfft(workers=%d) composition of the same chain, %.1f s" % (npasses, threads, ndt)}
        if cfg_times:
            line["configs"] = configs_report(cfg_times, cfg_ok, cfg_infos, world, peak)
        if world > 1 and torch.cuda.device_count() > 1:
            # handles on two GPUs inside ONE process (per-device kernel configuration, device switching inside the
            # library): the same frames through device 0 and device 1 must give the same bits
            try:
                chk = pcm[:4096].cpu().numpy()
                res = []
                for d in (0, 1, 0):
                    hd = usc.Handle(device=d)
                    res.append(hd.demod_frames_host(chk))
                    hd.close()
                line["two_devices_in_one_process"] = bool(all(np.array_equal(a, b) for r in res[1:] for a, b in zip(res[0], r)))
            except Exception as exc:
                line["two_devices_in_one_process"] = "error: %r" % (exc,)
        print(json.dumps(line))


if __name__ == "__main__":
    main()
