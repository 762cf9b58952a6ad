"""ctypes binding of the CPU oracle (oracle/libref_dsp.so).  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this module; the product (libusc.so and its Python mirror) never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libref_dsp.so")

f32p = C.POINTER(C.c_float)
u32p = C.POINTER(C.c_uint32)
i32p = C.POINTER(C.c_int32)

REF_MAX_RADICES = 8


class FftPlan(C.Structure):
    _fields_ = [("n", C.c_uint32), ("nrad", C.c_uint32), ("rad", C.c_uint32 * REF_MAX_RADICES),
                ("tw", f32p), ("tw_n", C.c_uint32)]


class RfftInstance(C.Structure):
    _fields_ = [("fftLenRFFT", C.c_uint16), ("cplx", FftPlan)]


class CfftInstance(C.Structure):
    _fields_ = [("fftLen", C.c_uint16), ("plan", FftPlan)]


class FirInstance(C.Structure):
    _fields_ = [("numTaps", C.c_uint16), ("pState", f32p), ("pCoeffs", f32p)]


class ChirpParams(C.Structure):
    _fields_ = [("n", C.c_uint32), ("fs", C.c_float), ("f0", C.c_float), ("f1", C.c_float),
                ("sweep_T", C.c_float), ("phase", C.c_float)]


class History(C.Structure):
    _fields_ = [("mag_max", C.c_float), ("mag_max_left", C.c_float), ("mag_max_right", C.c_float),
                ("max_freq", C.c_int32), ("max_freq_left", C.c_int32), ("max_freq_right", C.c_int32),
                ("max_idx", C.c_uint32), ("max_idx_left", C.c_uint32), ("max_idx_right", C.c_uint32),
                ("mag_mean", C.c_float), ("snr", C.c_float), ("rank", C.c_char)]


class Receiver(C.Structure):
    _fields_ = [("n", C.c_uint32), ("fs", C.c_float), ("bandwidth", C.c_uint32),
                ("bandwidth2", C.c_uint32), ("idx_left_zero", C.c_uint32), ("hann", f32p),
                ("up_chirp", f32p), ("down_chirp", f32p), ("S", RfftInstance)]


class Compressor(C.Structure):
    _fields_ = [("n", C.c_uint32), ("fs", C.c_float), ("window", f32p), ("H_up", f32p),
                ("H_down", f32p), ("S", RfftInstance)]


def build(force=False):
    """Compile the oracle (gcc, seconds).  Building the checker is not using it."""
    if force or not os.path.exists(_LIB_PATH) or \
            os.path.getmtime(_LIB_PATH) < max(os.path.getmtime(os.path.join(_HERE, f))
                                              for f in ("ref_dsp.c", "ref_dsp.h", "ref_fast.c", "ref_fast_dispatch.c", "Makefile")):
        subprocess.check_call(["make", "-C", _HERE, "-s"], env=dict(os.environ, CC="gcc"))
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        L.ref_arm_cos_f32.restype = C.c_float
        L.ref_arm_cos_f32.argtypes = [C.c_float]
        L.ref_idx2freq.restype = C.c_int32
        L.ref_fft_radices.restype = C.c_uint32
        _lib = L
    return _lib


def _fp(a):
    return a.ctypes.data_as(f32p)


def _up(a):
    return a.ctypes.data_as(u32p)


def f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


# ---- primitives -------------------------------------------------------------------------------
def arm_cos_f32(x):
    L = lib()
    x = np.atleast_1d(np.asarray(x, dtype=np.float32))
    return np.array([L.ref_arm_cos_f32(C.c_float(float(v))) for v in x], dtype=np.float32)


def arm_sin_cos_f32(theta_deg):
    s, c = C.c_float(), C.c_float()
    lib().ref_arm_sin_cos_f32(C.c_float(float(theta_deg)), C.byref(s), C.byref(c))
    return np.float32(s.value), np.float32(c.value)


def hann_window(n, symmetric=False):
    w = np.empty(n, np.float32)
    lib().ref_hann_window(_fp(w), C.c_uint32(n), C.c_int(1 if symmetric else 0))
    return w


def generate_ref_chirp(variant, n, fs, f0, f1, sweep_T, phase, up):
    v = {"R": 0, "S": 1, "T": 2, "F": 3}[variant]
    out = np.empty(2 * n if variant == "S" else n, np.float32)
    p = ChirpParams(n, fs, f0, f1, sweep_T, phase)
    lib().ref_generate_ref_chirp(C.c_int(v), C.byref(p), C.c_int(1 if up else 0), _fp(out))
    return out


def arm_mult_f32(a, b):
    a, b = f32(a), f32(b)
    d = np.empty_like(a)
    lib().ref_arm_mult_f32(_fp(a), _fp(b), _fp(d), C.c_uint32(a.size))
    return d


def arm_scale_f32(a, s):
    a = f32(a)
    d = np.empty_like(a)
    lib().ref_arm_scale_f32(_fp(a), C.c_float(s), _fp(d), C.c_uint32(a.size))
    return d


def arm_mean_f32(a):
    a = f32(a)
    r = C.c_float()
    lib().ref_arm_mean_f32(_fp(a), C.c_uint32(a.size), C.byref(r))
    return np.float32(r.value)


def arm_max_f32(a):
    a = f32(a)
    r, i = C.c_float(), C.c_uint32()
    lib().ref_arm_max_f32(_fp(a), C.c_uint32(a.size), C.byref(r), C.byref(i))
    return np.float32(r.value), int(i.value)


def arm_cmplx_mult_cmplx_f32(a, b):
    a, b = f32(a), f32(b)
    d = np.empty_like(a)
    lib().ref_arm_cmplx_mult_cmplx_f32(_fp(a), _fp(b), _fp(d), C.c_uint32(a.size // 2))
    return d


def arm_cmplx_mult_real_f32(a, r):
    a, r = f32(a), f32(r)
    d = np.empty_like(a)
    lib().ref_arm_cmplx_mult_real_f32(_fp(a), _fp(r), _fp(d), C.c_uint32(r.size))
    return d


def arm_cmplx_mag_f32(a):
    a = f32(a)
    d = np.empty(a.size // 2, np.float32)
    lib().ref_arm_cmplx_mag_f32(_fp(a), _fp(d), C.c_uint32(a.size // 2))
    return d


class Rfft:
    """arm_rfft_fast_instance_f32 + arm_rfft_fast_f32 (packed spectrum)."""

    def __init__(self, n):
        self.n = n
        self.S = RfftInstance()
        st = lib().ref_arm_rfft_fast_init_f32(C.byref(self.S), C.c_uint32(n))
        if st != 0:
            raise ValueError("ARM_MATH_ARGUMENT_ERROR")

    def __call__(self, x, inverse=False):
        x = f32(x).copy()
        out = np.empty(self.n, np.float32)
        lib().ref_arm_rfft_fast_f32(C.byref(self.S), _fp(x), _fp(out), C.c_uint8(1 if inverse else 0))
        return out

    def __del__(self):
        try:
            lib().ref_arm_rfft_fast_free(C.byref(self.S))
        except Exception:
            pass


class Cfft:
    """arm_cfft_instance_f32 + arm_cfft_f32 (in place, interleaved, natural order)."""

    def __init__(self, n):
        self.n = n
        self.S = CfftInstance()
        st = lib().ref_arm_cfft_init_f32(C.byref(self.S), C.c_uint32(n))
        if st != 0:
            raise ValueError("ARM_MATH_ARGUMENT_ERROR")

    def __call__(self, x, inverse=False):
        x = f32(x).copy()
        lib().ref_arm_cfft_f32(C.byref(self.S), _fp(x), C.c_uint8(1 if inverse else 0), C.c_uint8(1))
        return x

    def __del__(self):
        try:
            lib().ref_arm_cfft_free(C.byref(self.S))
        except Exception:
            pass


class Fir:
    def __init__(self, coeffs_reversed, block):
        self.c = f32(coeffs_reversed)
        self.block = block
        self.state = np.zeros(self.c.size + block - 1, np.float32)
        self.S = FirInstance()
        lib().ref_arm_fir_init_f32(C.byref(self.S), C.c_uint16(self.c.size), _fp(self.c),
                                   _fp(self.state), C.c_uint32(block))

    def __call__(self, x):
        x = f32(x)
        d = np.empty_like(x)
        lib().ref_arm_fir_f32(C.byref(self.S), _fp(x), _fp(d), C.c_uint32(x.size))
        return d


def twiddle(j, n):
    re, im = C.c_float(), C.c_float()
    lib().ref_twiddle(C.c_uint32(j), C.c_uint32(n), C.byref(re), C.byref(im))
    return np.float32(re.value), np.float32(im.value)


def fft_radices(n):
    rad = (C.c_uint32 * REF_MAX_RADICES)()
    c = lib().ref_fft_radices(C.c_uint32(n), rad)
    return [int(rad[i]) for i in range(c)]


# ---- receiver chain ---------------------------------------------------------------------------
class RefReceiver:
    """receiver/Src/main.c DSP chain (variant R tables)."""

    def __init__(self, n=2048, fs=78125.0, f0=16000.0, f1=19000.0, sweep_T=0.0205):
        self.rx = Receiver()
        if lib().ref_receiver_init(C.byref(self.rx), C.c_uint32(n), C.c_float(fs), C.c_float(f0),
                                   C.c_float(f1), C.c_float(sweep_T)) != 0:
            raise RuntimeError("ref_receiver_init failed")
        self.n = n

    @property
    def bandwidth2(self):
        return int(self.rx.bandwidth2)

    @property
    def idx_left_zero(self):
        return int(self.rx.idx_left_zero)

    def table(self, name):
        return np.ctypeslib.as_array(getattr(self.rx, name), shape=(self.n,)).copy()

    def idx2freq(self, idx):
        return int(lib().ref_idx2freq(C.byref(self.rx), C.c_uint32(idx)))

    def pipeline(self, signal, up):
        s = f32(signal).copy()
        lib().ref_pipeline(C.byref(self.rx), _fp(s), C.c_int(1 if up else 0))
        return s

    def dsp(self, fifo, sync_position, mag_mean, up):
        fifo = f32(fifo)
        h = History()
        lib().ref_dsp(C.byref(self.rx), _fp(fifo), C.c_uint32(sync_position), C.byref(h),
                      C.c_float(mag_mean), C.c_int(1 if up else 0))
        return h

    def demod_frames(self, pcm, nthreads=1):
        """pcm: (nframes, n) int32 or float32 -> (mag_up, idx_up, mag_down, idx_down)."""
        pcm = np.ascontiguousarray(pcm)
        nf = pcm.size // self.n
        mu, md = np.empty(nf, np.float32), np.empty(nf, np.float32)
        iu, idn = np.empty(nf, np.uint32), np.empty(nf, np.uint32)
        if pcm.dtype == np.int32:
            lib().ref_demod_frames_i32(C.byref(self.rx), pcm.ctypes.data_as(i32p), C.c_size_t(nf),
                                       _fp(mu), _up(iu), _fp(md), _up(idn), C.c_int(nthreads))
        elif pcm.dtype == np.float32:
            lib().ref_demod_frames_f32(C.byref(self.rx), _fp(pcm), C.c_size_t(nf),
                                       _fp(mu), _up(iu), _fp(md), _up(idn), C.c_int(nthreads))
        else:
            raise TypeError(pcm.dtype)
        return mu, iu, md, idn

    def demod_frames_fast(self, pcm, nthreads=1):
        """The tuned form (ref_fast.c): same results as demod_frames, bit for bit."""
        pcm = np.ascontiguousarray(pcm)
        if pcm.dtype not in (np.int32, np.float32):
            raise TypeError(pcm.dtype)
        nf = pcm.size // self.n
        mu, md = np.empty(nf, np.float32), np.empty(nf, np.float32)
        iu, idn = np.empty(nf, np.uint32), np.empty(nf, np.uint32)
        rc = lib().ref_fast_demod_frames(C.byref(self.rx), C.c_void_p(pcm.ctypes.data), C.c_int(pcm.dtype == np.float32),
                                         C.c_size_t(nf), _fp(mu), _up(iu), _fp(md), _up(idn), C.c_int(nthreads))
        if rc != 0:
            raise RuntimeError("ref_fast_demod_frames: unsupported receiver")
        return mu, iu, md, idn

    def __del__(self):
        try:
            lib().ref_receiver_free(C.byref(self.rx))
        except Exception:
            pass


class RefCompressor:
    """experiments/chirp_compression_time_domain chain."""

    def __init__(self, n=2048, fs=100000.0, f1=17000.0, f2=18000.0):
        self.c = Compressor()
        if lib().ref_compressor_init(C.byref(self.c), C.c_uint32(n), C.c_float(fs), C.c_float(f1),
                                     C.c_float(f2)) != 0:
            raise RuntimeError("ref_compressor_init failed")
        self.n = n

    def table(self, name):
        return np.ctypeslib.as_array(getattr(self.c, name), shape=(self.n,)).copy()

    def compress(self, frame, use_up=False):
        s = f32(frame).copy()
        lib().ref_compress_chirp(C.byref(self.c), _fp(s), C.c_int(1 if use_up else 0))
        return s

    def compress_frames(self, pcm, use_up=False, nthreads=1):
        pcm = np.ascontiguousarray(pcm, dtype=np.int32)
        nf = pcm.size // self.n
        mv, mi = np.empty(nf, np.float32), np.empty(nf, np.uint32)
        lib().ref_compress_frames_i32(C.byref(self.c), pcm.ctypes.data_as(i32p), C.c_size_t(nf),
                                      C.c_int(1 if use_up else 0), _fp(mv), _up(mi), C.c_int(nthreads))
        return mv, mi

    def __del__(self):
        try:
            lib().ref_compressor_free(C.byref(self.c))
        except Exception:
            pass


# ---- receiver state machine ---------------------------------------------------------------------
class RxState(C.Structure):
    _fields_ = [("state", C.c_uint32), ("turn", C.c_uint32), ("sync_cnt", C.c_uint32),
                ("sync_position", C.c_uint32), ("max_idx", C.c_uint32), ("mag_stat", C.c_float * 12),
                ("mag_mean", C.c_float), ("history", History * 8), ("msg", C.c_uint32), ("msg_cnt", C.c_uint32),
                ("lock_frame", C.c_int32), ("lock_position", C.c_uint32), ("frames_seen", C.c_uint32)]


def receiver_run(rxobj, pcm, snr_threshold=2.0, cap=4096):
    """Whole-stream state machine (receiver/Src/main.c:417-580): -> (uart bytes, final RxState)."""
    pcm = np.ascontiguousarray(pcm, dtype=np.int32)
    nframes = pcm.size // rxobj.n
    out = np.zeros(cap, np.uint8)
    nout = C.c_uint32(0)
    st = RxState()
    lib().ref_receiver_run_i32(C.byref(rxobj.rx), C.c_float(snr_threshold), pcm.ctypes.data_as(i32p),
                               C.c_uint32(nframes), out.ctypes.data_as(C.POINTER(C.c_uint8)), C.byref(nout),
                               C.c_uint32(cap), C.byref(st))
    return bytes(out[:min(nout.value, cap)]), st


def sync_search(rxobj, pcm, sync_add=1):
    pcm = np.ascontiguousarray(pcm, dtype=np.int32)
    nframes = pcm.size // rxobj.n
    mag = np.empty((nframes, 4), np.float32)
    idx = np.empty((nframes, 4), np.uint32)
    lib().ref_sync_search_i32(C.byref(rxobj.rx), pcm.ctypes.data_as(i32p), C.c_uint32(nframes), C.c_uint32(sync_add),
                              _fp(mag), _up(idx))
    return mag, idx


# ---- complex-FFT variant (experiments/synchronization) ---------------------------------------------
class SyncReceiver(C.Structure):
    _fields_ = [("n", C.c_uint32), ("fs", C.c_float), ("bandwidth", C.c_uint32), ("bandwidth2", C.c_uint32),
                ("idx_left_zero", C.c_uint32), ("hann", f32p), ("up_chirp", f32p), ("down_chirp", f32p),
                ("C", CfftInstance)]


class RefSyncReceiver:
    def __init__(self, n=2048, fs=78125.0, f0=16000.0, f1=19000.0, sweep_T=0.0205):
        self.rx = SyncReceiver()
        if lib().ref_sync_receiver_init(C.byref(self.rx), C.c_uint32(n), C.c_float(fs), C.c_float(f0), C.c_float(f1),
                                        C.c_float(sweep_T)) != 0:
            raise RuntimeError("ref_sync_receiver_init failed")
        self.n = n

    def table(self, name):
        ln = self.n if name == "hann" else 2 * self.n
        return np.ctypeslib.as_array(getattr(self.rx, name), shape=(ln,)).copy()

    def pipeline(self, cframe, up):
        s = f32(cframe).copy()
        lib().ref_sync_pipeline(C.byref(self.rx), _fp(s), C.c_int(1 if up else 0))
        return s[:self.n].copy()

    def dsp(self, fifo, sync_position, mag_mean, up):
        fifo = f32(fifo)
        h = History()
        lib().ref_sync_dsp(C.byref(self.rx), _fp(fifo), C.c_uint32(sync_position), C.byref(h), C.c_float(mag_mean),
                           C.c_int(1 if up else 0))
        return h

    def __del__(self):
        try:
            lib().ref_sync_receiver_free(C.byref(self.rx))
        except Exception:
            pass


# ---- I/Q baseband path ---------------------------------------------------------------------------
class Iq(C.Structure):
    _fields_ = [("n", C.c_uint32), ("fs", C.c_float), ("window_bins", C.c_uint32), ("carrier_cos", f32p),
                ("carrier_sin", f32p), ("chirp", f32p), ("chirp_conj", f32p), ("hann", f32p), ("taps", C.c_float * 64),
                ("num_taps", C.c_uint32), ("C", CfftInstance)]


class RefIq:
    def __init__(self, taps_reversed, n=2048, fs=78125.0, carrier=18000.0, bw=3000.0, sweep_T=0.0205, window_bins=32):
        self.q = Iq()
        t = f32(taps_reversed)
        if lib().ref_iq_init(C.byref(self.q), C.c_uint32(n), C.c_float(fs), C.c_float(carrier), C.c_float(bw),
                             C.c_float(sweep_T), _fp(t), C.c_uint32(t.size), C.c_uint32(window_bins)) != 0:
            raise RuntimeError("ref_iq_init failed")
        self.n = n

    def table(self, name):
        ln = {"carrier_cos": self.n, "carrier_sin": self.n, "chirp": self.n, "chirp_conj": self.n, "hann": self.n // 2}[name]
        return np.ctypeslib.as_array(getattr(self.q, name), shape=(ln,)).copy()

    def demod(self, pcm):
        """pcm: [nframes, n] int32 of ONE stream -> (mag_up, idx_up, mag_down, idx_down)."""
        pcm = np.ascontiguousarray(pcm, dtype=np.int32)
        nf = pcm.size // self.n
        mu, md = np.empty(nf, np.float32), np.empty(nf, np.float32)
        iu, idn = np.empty(nf, np.uint32), np.empty(nf, np.uint32)
        lib().ref_iq_demod_i32(C.byref(self.q), pcm.ctypes.data_as(i32p), C.c_uint32(nf), _fp(mu), _up(iu), _fp(md), _up(idn))
        return mu, iu, md, idn

    def __del__(self):
        try:
            lib().ref_iq_free(C.byref(self.q))
        except Exception:
            pass


def synth_frames(seed, first_frame, nframes, amp, noise_sigma, n=2048, fs=78125.0, f0=16000.0, f1=19000.0):
    """CPU twin of usc_synth_frames -> (pcm [nframes, n] int32, bits)."""
    pcm = np.empty((nframes, n), np.int32)
    bits = np.empty(nframes, np.uint8)
    lib().ref_synth_frames(C.c_uint64(seed), C.c_uint64(first_frame), C.c_size_t(nframes), C.c_uint32(n), C.c_float(fs),
                           C.c_float(f0), C.c_float(f1), C.c_double(amp), C.c_double(noise_sigma),
                           pcm.ctypes.data_as(i32p), bits.ctypes.data_as(C.POINTER(C.c_uint8)))
    return pcm, bits


def synth_iq_frames(seed, first_frame, nframes, carrier, bw, sideband, phase, amp, noise_sigma, n=2048, fs=78125.0):
    """CPU twin of usc_synth_iq_frames -> (pcm [nframes, n] int32, bits)."""
    pcm = np.empty((nframes, n), np.int32)
    bits = np.empty(nframes, np.uint8)
    lib().ref_synth_iq_frames(C.c_uint64(seed), C.c_uint64(first_frame), C.c_size_t(nframes), C.c_uint32(n), C.c_float(fs),
                              C.c_double(carrier), C.c_double(bw), C.c_int(sideband), C.c_double(phase), C.c_double(amp),
                              C.c_double(noise_sigma), pcm.ctypes.data_as(i32p), bits.ctypes.data_as(C.POINTER(C.c_uint8)))
    return pcm, bits


def correlate_os(tmpl, pcm):
    """twin of usc_correlate_os for one stream: pcm [nframes, n] int32 or float32 -> (out [nframes-1, n], max, idx)"""
    tmpl = f32(tmpl)
    n = tmpl.size
    pcm = np.ascontiguousarray(pcm)
    nf = pcm.size // n
    out = np.empty((nf - 1, n), np.float32)
    mv, mi = np.empty(nf - 1, np.float32), np.empty(nf - 1, np.uint32)
    pi = pcm.ctypes.data_as(i32p) if pcm.dtype == np.int32 else None
    pf = _fp(pcm) if pcm.dtype == np.float32 else None
    if pi is None and pf is None:
        raise TypeError(pcm.dtype)
    lib().ref_correlate_os(_fp(tmpl), C.c_uint32(n), pi, pf, C.c_uint32(nf), _fp(out), _fp(mv), _up(mi))
    return out, mv, mi


def legacy_magnitudes(pcm, nthreads=4):
    """[nframes, n] int32 -> [nframes, n/2] magnitudes of the legacy detectors' shared front half."""
    pcm = np.ascontiguousarray(pcm, dtype=np.int32)
    nf, n = pcm.shape
    mag = np.empty((nf, n // 2), np.float32)
    lib().ref_legacy_magnitudes_i32(pcm.ctypes.data_as(i32p), C.c_size_t(nf), C.c_uint32(n), _fp(mag), C.c_int(nthreads))
    return mag


def onoff_detect(mag, fs=78125.0, f1=17000.0, f2=18000.0, mag_threshold=3000.0, high_frac=0.1, low_frac=0.05):
    nf, half = mag.shape
    lo, hi = C.c_uint32(0), C.c_uint32(0)
    lib().ref_onoff_band(C.c_uint32(2 * half), C.c_float(fs), C.c_float(f1), C.c_float(f2), C.byref(lo), C.byref(hi))
    strength, level = np.empty(nf, np.uint16), np.empty(nf, np.int8)
    lib().ref_onoff_levels(_fp(mag), C.c_size_t(nf), C.c_uint32(half), lo, hi, C.c_float(mag_threshold), C.c_float(high_frac),
                           C.c_float(low_frac), strength.ctypes.data_as(C.POINTER(C.c_uint16)),
                           level.ctypes.data_as(C.POINTER(C.c_int8)))
    return strength, level, (lo.value, hi.value)


def onoff_decode(level, frame_start=3, frame_bit=2, sync_threshold=2, sampling_offset=1, cap=64):
    level = np.ascontiguousarray(level, dtype=np.int8)
    chars = np.zeros(cap, np.uint8)
    errs = C.c_uint32(0)
    lib().ref_onoff_decode.restype = C.c_uint32
    n = lib().ref_onoff_decode(level.ctypes.data_as(C.POINTER(C.c_int8)), C.c_uint32(level.size), C.c_uint32(frame_start),
                               C.c_uint32(frame_bit), C.c_uint32(sync_threshold), C.c_uint32(sampling_offset),
                               chars.ctypes.data_as(C.POINTER(C.c_uint8)), C.c_uint32(cap), C.byref(errs))
    return bytes(chars[:min(n, cap)]), n, errs.value


def fsk_codes(mag, fs=78125.0, sof_bin=340, eof_bin=344, hex0_bin=348, hex_step=4, tolerance=0, mag_threshold=5000.0):
    nf, half = mag.shape
    code, m, fr = np.empty(nf, np.uint8), np.empty(nf, np.float32), np.empty(nf, np.float32)
    lib().ref_fsk_codes(_fp(mag), C.c_size_t(nf), C.c_uint32(half), C.c_float(fs), C.c_uint32(2 * half), C.c_uint32(sof_bin),
                        C.c_uint32(eof_bin), C.c_uint32(hex0_bin), C.c_uint32(hex_step), C.c_uint32(tolerance),
                        C.c_float(mag_threshold), code.ctypes.data_as(C.POINTER(C.c_uint8)), _fp(m), _fp(fr))
    return code, m, fr


def fsk_parse(code, tq_n=2, cap=64):
    code = np.ascontiguousarray(code, dtype=np.uint8)
    chars = np.zeros(cap, np.uint8)
    sof, eof = C.c_uint32(0), C.c_uint32(0)
    lib().ref_fsk_parse.restype = C.c_uint32
    n = lib().ref_fsk_parse(code.ctypes.data_as(C.POINTER(C.c_uint8)), C.c_uint32(code.size), C.c_uint32(tq_n),
                            chars.ctypes.data_as(C.POINTER(C.c_uint8)), C.c_uint32(cap), C.byref(sof), C.byref(eof))
    return bytes(chars[:min(n, cap)]), n, sof.value, eof.value


def resample_i16_to_pcm(x, up, down, n_out=None):
    x = np.ascontiguousarray(x, dtype=np.int16)
    if n_out is None:
        n_out = (x.size * up + down - 1) // down
    out = np.empty(n_out, np.int32)
    lib().ref_resample_i16_to_pcm(x.ctypes.data_as(C.c_void_p), C.c_size_t(x.size), C.c_uint32(up), C.c_uint32(down),
                                  out.ctypes.data_as(i32p), C.c_size_t(n_out))
    return out


def synth_streams(seed, first_stream, nstreams, nframes, lead_in, msg_bytes, guard, amp, noise_sigma, n=2048, fs=78125.0,
                  f0=16000.0, f1=19000.0):
    """CPU twin of usc_synth_streams -> (pcm [nstreams, nframes, n] int32, offsets, messages [nstreams, msg_bytes])."""
    pcm = np.empty((nstreams, nframes, n), np.int32)
    offs = np.empty(nstreams, np.uint32)
    msgs = np.empty((nstreams, max(msg_bytes, 1)), np.uint8)
    lib().ref_synth_streams(C.c_uint64(seed), C.c_uint64(first_stream), C.c_uint32(nstreams), C.c_uint32(nframes), C.c_uint32(n),
                            C.c_float(fs), C.c_float(f0), C.c_float(f1), C.c_uint32(lead_in), C.c_uint32(msg_bytes),
                            C.c_uint32(guard), C.c_double(amp), C.c_double(noise_sigma), pcm.ctypes.data_as(i32p), _up(offs),
                            msgs.ctypes.data_as(C.POINTER(C.c_uint8)))
    return pcm, offs, msgs[:, :msg_bytes]


class ScanEntry(C.Structure):
    _fields_ = [("mag_max_right", C.c_float), ("mag_max_left", C.c_float), ("max_idx_right", C.c_uint32),
                ("max_idx_left", C.c_uint32)]


def scan4(pcm2n, n=2048, fs=100000.0, f1=17000.0, f2=18000.0):
    """experiments/chirp_compression_freq_domain 4-offset scan on one 2n-sample buffer -> 4 ScanEntry."""
    r = Rfft(n)
    hann = hann_window(n)
    chirp = generate_ref_chirp("F", n, fs, f1, f2, 0.0, 0.0, 0)
    bandwidth = int(np.float32((int(f2 - f1) * n)) / np.float32(fs))
    buf = f32(pcm2n)
    out = (ScanEntry * 4)()
    lib().ref_scan4(C.byref(r.S), _fp(hann), _fp(chirp), C.c_uint32(n), C.c_uint32(bandwidth), _fp(buf), out)
    return [(e.mag_max_right, e.max_idx_right, e.mag_max_left, e.max_idx_left) for e in out]
