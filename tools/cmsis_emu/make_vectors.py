#!/usr/bin/env python
"""Golden vectors from the reference's OWN CMSIS-DSP binary (container only: needs /root/reference).

The reference vendors CMSIS-DSP V1.4.5b as receiver/Drivers/CMSIS/Lib/libarm_cortexM4lf_math.a (GCC 5.4, ARM Thumb-2,
hard-float).  tools/cmsis_emu/thumb2.py links the needed members and interprets their machine code, so every number
written here was produced by the reference's real arithmetic — the chirp tables by its arm_sin_cos_f32 / arm_cos_f32,
the spectra by its arm_rfft_fast_f32 / arm_cfft_f32, and so on.  The glue between the calls (which function, which
buffer, which length) follows the reference's C sources and cites them.

Output: tests/golden/cmsis_binary_vectors.npz  (read by tests/test_cmsis_binary.py and tests/test_gpu_cmsis_binary.py)

    python tools/cmsis_emu/make_vectors.py            # about five minutes
"""
import os
import struct
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)
from thumb2 import Cpu, Linker, read_archive  # noqa: E402

ARCHIVE = "/root/reference/receiver/Drivers/CMSIS/Lib/libarm_cortexM4lf_math.a"
MEMBERS = ["arm_sin_cos_f32.o", "arm_cos_f32.o", "arm_sin_f32.o", "arm_common_tables.o", "arm_const_structs.o",
           "arm_rfft_fast_init_f32.o", "arm_rfft_fast_f32.o", "arm_cfft_f32.o", "arm_cfft_radix8_f32.o",
           "arm_bitreversal2.o", "arm_cmplx_mag_f32.o", "arm_cmplx_mult_cmplx_f32.o", "arm_cmplx_mult_real_f32.o",
           "arm_mult_f32.o", "arm_scale_f32.o", "arm_max_f32.o", "arm_mean_f32.o", "arm_fir_init_f32.o", "arm_fir_f32.o",
           "arm_copy_f32.o"]
f32 = np.float32


class Cmsis:
    """The archive's functions on numpy arrays (every call runs the reference's machine code)."""

    def __init__(self):
        ar = read_archive(ARCHIVE)
        self.cpu = Cpu(1 << 24)
        self.ln = Linker(self.cpu)

        def memset(cpu):                                          # the only libc routine these members call
            a, v, n = cpu.r[0], cpu.r[1] & 0xFF, cpu.r[2]
            cpu.mem[a:a + n] = bytes([v]) * n
        self.ln.add_hook("memset", memset)
        for m in MEMBERS:
            if m in ar:
                self.ln.add_object(ar[m])
        self.ln.resolve()
        self.heap0 = self.ln.alloc(0, 64)
        self.rfft_inst = {}

    def fn(self, name):
        return self.ln.symbols[name]

    def put(self, arr):
        raw = np.ascontiguousarray(arr).tobytes()
        a = self.ln.alloc(len(raw), 16)
        self.cpu.mem[a:a + len(raw)] = raw
        return a

    def get(self, a, n, dtype=np.float32):
        return np.frombuffer(bytes(self.cpu.mem[a:a + n * np.dtype(dtype).itemsize]), dtype=dtype).copy()

    def call(self, name, *args, s0=None):
        sp = len(self.cpu.mem) - 256
        for k, extra in enumerate(args[4:]):                      # AAPCS: arguments beyond r0-r3 go on the stack
            self.cpu.wr32(sp + 4 * k, extra)
        args = args[:4]
        self.cpu.r[13] = sp
        if s0 is not None:
            self.cpu.s[0] = struct.unpack("<I", f32(s0).tobytes())[0]
        return self.cpu.call(self.fn(name), args)

    def scratch(self):
        class _Scope:
            def __enter__(s):
                s.mark = self.ln.cursor
            def __exit__(s, *a):
                self.ln.cursor = s.mark
        return _Scope()

    # scalar functions
    def sin_cos(self, theta):
        with self.scratch():
            a = self.put(np.zeros(2, f32))
            self.call("arm_sin_cos_f32", a, a + 4, s0=theta)
            r = self.get(a, 2)
        return r[0], r[1]

    def cos(self, x):
        self.call("arm_cos_f32", s0=x)
        return np.frombuffer(struct.pack("<I", self.cpu.s[0]), dtype=f32)[0]

    # vector functions
    def binary(self, name, a, b, n_out, count):
        with self.scratch():
            pa, pb = self.put(f32(a)), self.put(f32(b))
            pd = self.put(np.zeros(n_out, f32))
            self.call(name, pa, pb, pd, count)
            return self.get(pd, n_out)

    def mult(self, a, b):
        return self.binary("arm_mult_f32", a, b, len(a), len(a))

    def cmul(self, a, b):
        return self.binary("arm_cmplx_mult_cmplx_f32", a, b, len(a), len(a) // 2)

    def cmul_real(self, a, r):
        return self.binary("arm_cmplx_mult_real_f32", a, r, len(a), len(r))

    def mag(self, a):
        with self.scratch():
            pa = self.put(f32(a))
            pd = self.put(np.zeros(len(a) // 2, f32))
            self.call("arm_cmplx_mag_f32", pa, pd, len(a) // 2)
            return self.get(pd, len(a) // 2)

    def scale(self, a, s):
        with self.scratch():
            pa = self.put(f32(a))
            pd = self.put(np.zeros(len(a), f32))
            self.call("arm_scale_f32", pa, pd, len(a), s0=s)       # hard-float ABI: (pSrc, scale in s0, pDst, blockSize)
            return self.get(pd, len(a))

    def max(self, a):
        with self.scratch():
            pa = self.put(f32(a))
            pr = self.put(np.zeros(2, f32))
            self.call("arm_max_f32", pa, len(a), pr, pr + 4)
            return self.get(pr, 1)[0], int(self.get(pr + 4, 1, np.uint32)[0])

    def mean(self, a):
        with self.scratch():
            pa = self.put(f32(a))
            pr = self.put(np.zeros(1, f32))
            self.call("arm_mean_f32", pa, len(a), pr)
            return self.get(pr, 1)[0]

    def rfft(self, x, inverse=False):
        n = len(x)
        if n not in self.rfft_inst:
            inst = self.ln.alloc(64, 8)
            self.heap0 = max(self.heap0, self.ln.cursor)
            assert self.call("arm_rfft_fast_init_f32", inst, n) == 0
            self.rfft_inst[n] = inst
        with self.scratch():
            pa = self.put(f32(x))
            pd = self.put(np.zeros(n, f32))
            self.call("arm_rfft_fast_f32", self.rfft_inst[n], pa, pd, 1 if inverse else 0)
            return self.get(pd, n)

    def cfft(self, z, inverse=False):
        """z: interleaved complex, in place, bit reversal on (how every reference call site uses it)"""
        n = len(z) // 2
        with self.scratch():
            pa = self.put(f32(z))
            self.call("arm_cfft_f32", self.fn("arm_cfft_sR_f32_len%d" % n), pa, 1 if inverse else 0, 1)
            return self.get(pa, 2 * n)

    def fir_new(self, taps, block):
        taps = f32(taps)
        inst = self.ln.alloc(16, 8)
        coef = self.put(taps)
        state = self.put(np.zeros(len(taps) + block - 1 + 8, f32))
        self.heap0 = max(self.heap0, self.ln.cursor)
        self.call("arm_fir_init_f32", inst, len(taps), coef, state, block)
        return inst

    def fir(self, inst, x):
        with self.scratch():
            pa = self.put(f32(x))
            pd = self.put(np.zeros(len(x), f32))
            self.call("arm_fir_f32", inst, pa, pd, len(x))
            return self.get(pd, len(x))


# ---- table builders: the reference's C glue restated with its expression types, CMSIS calls on the binary -----------
def receiver_chirp(cm, n, fs, f0, f1, sweep_t, phase, up, both=False):
    """receiver/Src/chirp.c:16-40 (both=False: the sine overwrites the cosine) and experiments/synchronization/
    Src/chirp.c:17-45 (both=True: interleaved cos, sin)."""
    out = np.empty(2 * n if both else n, f32)
    t = f32(0.0)
    delta_f = f32(f32(f1 - f0) / f32(sweep_t))
    delta_t = f32(f32(sweep_t) / f32(f32(sweep_t) * f32(fs)))
    for i in range(n):
        half = float(f32(delta_f * t)) / 2.0                                   # `delta_f * t / 2.0`: double division
        freq = f32(float(f0) + half) if up else f32(float(f1) - half)
        theta = f32(360.0 * float(freq) * float(t) + float(f32(phase)))
        t = f32(t + delta_t)
        s, c = cm.sin_cos(theta)
        if both:
            out[2 * i], out[2 * i + 1] = c, s
        else:
            out[i] = s
    return out


def hann_periodic(cm, n):
    """receiver/Src/main.c:99,390-393: WINDOW_SCALE = 2.0f * M_PI / (float) NN (double expression stored to float)"""
    scale = f32(2.0 * np.pi / float(n))
    return np.array([f32(0.5) - f32(f32(0.5) * cm.cos(f32(f32(i) * scale))) for i in range(n)], f32)


def hann_symmetric(cm, n):
    """experiments/chirp_compression_time_domain/Src/chirp.c:13,63-65: 2.0f * PI / (float)(PCM_SAMPLES - 1), float PI"""
    scale = f32(f32(f32(2.0) * f32(3.14159265358979)) / f32(n - 1))
    return np.array([f32(0.5) - f32(f32(0.5) * cm.cos(f32(f32(i) * scale))) for i in range(n)], f32)


def compression_chirp(cm, n, fs, f1, f2, phase, up):
    """experiments/chirp_compression_time_domain/Src/chirp.c:25-45: radians, arm_cos_f32, float slope without /2"""
    out = np.empty(n, f32)
    t = f32(0.0)
    time_frame = f32(f32(n) / f32(fs))
    delta_f = f32(f32(f32(f2) - f32(f1)) / time_frame)
    delta_t = f32(time_frame / f32(time_frame * f32(fs)))
    two_pi = 2.0 * float(f32(3.14159265358979))
    for i in range(n):
        freq = f32(f32(f1) + f32(delta_f * t)) if up else f32(f32(f2) - f32(delta_f * t))
        arg = f32(two_pi * float(freq) * float(t) + float(f32(phase)))
        t = f32(t + delta_t)
        out[i] = cm.cos(arg)
    return out


def main():
    from oracle import pyref
    t0 = time.time()
    cm = Cmsis()
    out = {}
    rng = np.random.default_rng(20261017)

    # 1. tables as the archive stores / computes them
    tab = cm.fn("sinTable_f32")
    out["sinTable_f32"] = cm.get(tab, 513)
    thetas = np.concatenate([rng.uniform(-400000, 400000, 300), rng.uniform(-360, 360, 100), [0.0, 90.0, -90.0, 180.0, 359.99]]).astype(f32)
    sc = np.array([cm.sin_cos(t) for t in thetas], f32)
    out["sin_cos_theta"], out["sin_cos_sin"], out["sin_cos_cos"] = thetas, sc[:, 0], sc[:, 1]
    xs = rng.uniform(-60, 60, 400).astype(f32)
    out["cos_x"], out["cos_y"] = xs, np.array([cm.cos(x) for x in xs], f32)
    print("tables", time.time() - t0)

    # 2. receiver chain (receiver/Src/main.c:163-215) on frames of the bench dataset and cleaner ones
    N, FS, F0, F1, SWEEP = 2048, 78125.0, 16000.0, 19000.0, 0.0205
    up = receiver_chirp(cm, N, FS, F0, F1, SWEEP, -90.0, True)
    down = receiver_chirp(cm, N, FS, F0, F1, SWEEP, -90.0, False)
    hann = hann_periodic(cm, N)
    out["rx_up_chirp"], out["rx_down_chirp"], out["rx_hann"] = up, down, hann
    bw2 = 2 * int(f32(f32((int(F1 - F0)) * N) / f32(FS)))                      # main.c:372-373: bandwidth2 = bandwidth * 2
    cases = [(7, 0, 20, 2.0e4, 2.0e4 / (10 ** (-5.0 / 20))),                    # bench.py's config-2 dataset, first frames
             (11, 0, 6, 2.0e4, 0.0), (12, 5, 6, 2.0e4, 2.0e4)]
    seeds, pcm_all = [], []
    for seed, first, nf, amp, sigma in cases:
        pcm, _ = pyref.synth_frames(seed, first, nf, amp, sigma)
        pcm_all.append(pcm)
        seeds.append((seed, first, nf, amp, sigma))
    pcm = np.concatenate(pcm_all)
    out["rx_cases"] = np.array(seeds, np.float64)
    out["rx_pcm_crc"] = np.array([int(np.bitwise_xor.reduce(pcm.view(np.uint32).ravel()))], np.uint32)
    mags = np.empty((len(pcm), 2, N // 2), f32)
    peak = np.empty((len(pcm), 2), f32)
    pidx = np.empty((len(pcm), 2), np.uint32)
    for f, frame in enumerate(pcm):
        x = frame.astype(f32)                                                    # main.c:663-665
        for h, chirp in enumerate((up, down)):
            sig = cm.mult(x, chirp)                                              # chirp.c:47-53
            sig = cm.mult(sig, hann)                                             # main.c:171
            spec = cm.rfft(sig)                                                  # main.c:174
            m = cm.mag(spec)                                                     # main.c:178 (the valid half)
            mags[f, h] = m
            peak[f, h], pidx[f, h] = cm.max(m[:bw2])                             # main.c:208
        if f % 8 == 0:
            print("rx frame", f, time.time() - t0)
    out["rx_mag"], out["rx_peak"], out["rx_peak_idx"], out["rx_bw2"] = mags, peak, pidx, np.array([bw2], np.uint32)

    # 3. time-domain compression chain (experiments/chirp_compression_time_domain/Src/chirp.c:50-83, main.c:189)
    CFS, CF1, CF2 = 100000.0, 17000.0, 18000.0
    win = hann_symmetric(cm, N)
    h_tabs = {}
    for name, upf in (("up", True), ("down", False)):
        c = compression_chirp(cm, N, CFS, CF1, CF2, f32(-3.14159265358979 / 2.0), upf)
        h_tabs[name] = cm.rfft(cm.mult(c, win))
        out["cc_chirp_" + name] = c
        out["cc_H_" + name] = h_tabs[name]
    out["cc_hann"] = win
    cpcm, _ = pyref.synth_frames(21, 0, 6, 2.0e4, 1.0e4, n=N, fs=CFS, f0=CF1, f1=CF2)
    out["cc_case"] = np.array([21, 0, 6, 2.0e4, 1.0e4], np.float64)
    comp = np.empty((len(cpcm), N), f32)
    cmax = np.empty(len(cpcm), f32)
    cidx = np.empty(len(cpcm), np.uint32)
    for f, frame in enumerate(cpcm):
        sig = cm.mult(frame.astype(f32), win)
        spec = cm.rfft(sig)
        prod = cm.cmul(spec, h_tabs["down"])
        comp[f] = cm.rfft(prod, inverse=True)
        cmax[f], cidx[f] = cm.max(comp[f])
    out["cc_out"], out["cc_max"], out["cc_idx"] = comp, cmax, cidx
    print("compress", time.time() - t0)

    # 4. complex transforms: arm_cfft_f32 at 1024 (I/Q back end) and 2048 (experiments/synchronization)
    for n in (1024, 2048):
        z = rng.standard_normal(2 * n).astype(f32)
        out["cfft%d_in" % n] = z
        out["cfft%d_out" % n] = cm.cfft(z)
        out["cifft%d_out" % n] = cm.cfft(z, inverse=True)
    # real transforms at the other lengths the handle accepts that CMSIS has too
    for n in (256, 1024, 4096):
        x = rng.standard_normal(n).astype(f32)
        out["rfft%d_in" % n] = x
        out["rfft%d_out" % n] = cm.rfft(x)
        out["rifft%d_out" % n] = cm.rfft(out["rfft%d_out" % n], inverse=True)
    print("fft", time.time() - t0)

    # 5. I/Q front end (experiments/iq_modulation/Src/iq_modem.c:34-66): carrier tables, mix, two FIRs with carried state
    taps = np.load(os.path.join(ROOT, "tests", "golden", "fir_taps.npz"))["taps"].astype(f32)
    t = f32(0.0)
    delta_t = f32(f32(SWEEP) / f32(f32(SWEEP) * f32(FS)))
    csin, ccos = np.empty(N, f32), np.empty(N, f32)
    for i in range(N):
        theta = f32(360.0 * 18000.0 * float(t))
        csin[i], ccos[i] = cm.sin_cos(theta)
        t = f32(t + delta_t)
    out["iq_carrier_sin"], out["iq_carrier_cos"] = csin, ccos
    fi, fq = cm.fir_new(taps, N), cm.fir_new(taps, N)
    ipcm, _ = pyref.synth_iq_frames(31, 0, 3, 18000.0, 3000.0, -1, 0.0, 2.0e4, 6000.0)
    out["iq_case"] = np.array([31, 0, 3, 18000.0, 3000.0, -1, 0.0, 2.0e4, 6000.0], np.float64)
    iq_i, iq_q = np.empty((3, N), f32), np.empty((3, N), f32)
    for f, frame in enumerate(ipcm):
        x = frame.astype(f32)
        q = cm.mult(x, csin)                                                     # iq_modem.c:55
        i_ = cm.mult(x, ccos)                                                    # iq_modem.c:56
        iq_i[f] = cm.fir(fi, i_)                                                 # iq_modem.c:59
        iq_q[f] = cm.fir(fq, q)                                                  # iq_modem.c:60
    out["iq_fir_i"], out["iq_fir_q"], out["iq_taps"] = iq_i, iq_q, taps
    print("iq", time.time() - t0)

    # 6. the small operators
    a, b = rng.standard_normal(512).astype(f32), rng.standard_normal(512).astype(f32)
    out["op_a"], out["op_b"] = a, b
    out["op_mult"] = cm.mult(a, b)
    out["op_cmul"] = cm.cmul(a, b)
    out["op_cmul_real"] = cm.cmul_real(a, b[:256])
    out["op_mag"] = cm.mag(a)
    out["op_scale"] = cm.scale(a, f32(0.022097087))
    mx = a.copy()
    mx[100] = mx[300] = f32(9.5)                                                 # a tie: the first index wins
    out["op_max_in"] = mx
    v, i = cm.max(mx)
    out["op_max"], out["op_max_idx"] = np.array([v], f32), np.array([i], np.uint32)
    out["op_mean"] = np.array([cm.mean(a), cm.mean(a[:37])], f32)

    out["provenance"] = np.array(["executed from " + ARCHIVE + " by tools/cmsis_emu (thumb2.py); %d instructions" % cm.cpu.icount])
    path = os.path.join(ROOT, "tests", "golden", "cmsis_binary_vectors.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes;", cm.cpu.icount, "instructions;", time.time() - t0, "s")
    if cm.ln.undefined:
        print("symbols left undefined (never reached):", sorted(cm.ln.undefined)[:12])


if __name__ == "__main__":
    main()
