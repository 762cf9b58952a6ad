"""usc — Python mirror of the libusc.so C-ABI (include/usc.h), used by the tests and bench.py.

Product code: it binds ONLY libusc.so (hand-written sm_100a kernels behind a C ABI).  There is no
CPU fallback: if the library is missing or no CUDA device is usable, construction raises.
Device buffers are plain integers (device addresses) or anything exposing `.data_ptr()` (torch
tensors) / `__cuda_array_interface__`.
"""
import ctypes as C
import os

import numpy as np

_PKG = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB_PATH = os.environ.get("USC_LIB") or os.path.join(_PKG, "libusc.so")   # USC_LIB: A/B a build of the same ABI

USC_OK = 0
USC_ERR_ARGUMENT = -1
CHIRP_R, CHIRP_S, CHIRP_T, CHIRP_F = 0, 1, 2, 3
HANN_PERIODIC, HANN_SYMMETRIC = 0, 1
PCM_F32, PCM_I32 = 0, 1
UP, DOWN = 1, 0


class Config(C.Structure):
    _fields_ = [("n", C.c_uint32), ("fs", C.c_float), ("f0", C.c_float), ("f1", C.c_float),
                ("sweep_T", C.c_float), ("chirp_variant", C.c_uint32), ("window", C.c_uint32),
                ("snr_threshold", C.c_float), ("reserved", C.c_uint32 * 4)]


history_dtype = np.dtype([("mag_max", "f4"), ("mag_max_left", "f4"), ("mag_max_right", "f4"),
                          ("max_freq", "i4"), ("max_freq_left", "i4"), ("max_freq_right", "i4"),
                          ("max_idx", "u4"), ("max_idx_left", "u4"), ("max_idx_right", "u4"),
                          ("mag_mean", "f4"), ("snr", "f4"), ("rank", "u4")])

rx_result_dtype = np.dtype([("state", "u4"), ("sync_position", "u4"), ("lock_frame", "i4"), ("lock_position", "u4"),
                            ("nbytes", "u4"), ("frames_seen", "u4"), ("turn", "u4"), ("sync_cnt", "u4")])

scan_entry_dtype = np.dtype([("mag_max_right", "f4"), ("mag_max_left", "f4"), ("max_idx_right", "u4"), ("max_idx_left", "u4")])

class OnOffConfig(C.Structure):
    _fields_ = [("f1_hz", C.c_float), ("f2_hz", C.c_float), ("magnitude_threshold", C.c_float), ("high_frac", C.c_float),
                ("low_frac", C.c_float), ("frame_start", C.c_uint32), ("frame_bit", C.c_uint32), ("sync_threshold", C.c_uint32),
                ("sampling_offset", C.c_uint32)]


class FskConfig(C.Structure):
    _fields_ = [("sof_bin", C.c_uint32), ("eof_bin", C.c_uint32), ("hex0_bin", C.c_uint32), ("hex_step", C.c_uint32),
                ("tolerance", C.c_uint32), ("tq_n", C.c_uint32), ("magnitude_threshold", C.c_float)]


# every symbol include/usc.h declares (tests/test_abi.py checks the .so exports each one)
SYMBOLS = [
    "usc_default_config", "usc_create", "usc_destroy", "usc_set_stream", "usc_sync", "usc_error_string",
    "usc_get_geometry", "usc_get_table", "usc_launch_count", "usc_malloc", "usc_malloc_on", "usc_free", "usc_malloc_host",
    "usc_free_host", "usc_memcpy_h2d", "usc_memcpy_d2h", "usc_memset", "usc_i32_to_f32",
    "usc_arm_mult_f32_batch", "usc_arm_scale_f32_batch", "usc_arm_cmplx_mult_cmplx_f32_batch",
    "usc_arm_cmplx_mult_real_f32_batch", "usc_arm_cmplx_mag_f32_batch", "usc_arm_max_f32_batch",
    "usc_arm_mean_f32_batch", "usc_arm_rfft_fast_f32_batch", "usc_arm_cfft_f32_batch",
    "usc_arm_fir_f32_batch", "usc_demod_frames", "usc_host_workspace", "usc_demod_frames_host", "usc_iq_demod_host", "usc_receiver_run_host", "usc_receiver_run", "usc_receiver_run_chunk", "usc_sync_search", "usc_iq_init", "usc_iq_demod", "usc_spectrum_analyzer", "usc_synth_frames", "usc_synth_iq_frames", "usc_synth_streams", "usc_resample_i16_to_pcm", "usc_onoff_default_config", "usc_onoff_detect", "usc_fsk_default_config", "usc_fsk_detect", "usc_band_magnitudes", "usc_scan4", "usc_pipeline", "usc_dsp", "usc_compress_chirp", "usc_correlate_os",
]

_lib = None


def load():
    """dlopen libusc.so; raises (never falls back) if it is missing."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError("libusc.so not built: run `python ultrasonic-communication_b200/build.py` "
                               "(there is no CPU fallback)")
        L = C.CDLL(LIB_PATH)
        L.usc_error_string.restype = C.c_char_p
        L.usc_launch_count.restype = C.c_uint64
        L.usc_launch_count.argtypes = [C.c_void_p]
        if not os.environ.get("USC_LIB"):          # an A/B build may predate newer entry points
            for name in SYMBOLS:
                getattr(L, name)
        _lib = L
    return _lib


class UscError(RuntimeError):
    def __init__(self, code):
        self.code = code
        super().__init__("usc error %d: %s" % (code, load().usc_error_string(code).decode()))


def _ck(rc):
    if rc != 0:
        raise UscError(rc)


def _ptr(x):
    if x is None:
        return C.c_void_p(0)
    if isinstance(x, int):
        return C.c_void_p(x)
    if hasattr(x, "data_ptr"):
        return C.c_void_p(x.data_ptr())
    if hasattr(x, "__cuda_array_interface__"):
        return C.c_void_p(x.__cuda_array_interface__["data"][0])
    if hasattr(x, "ptr"):
        return C.c_void_p(x.ptr)
    raise TypeError("not a device buffer: %r" % (type(x),))


class DeviceBuffer:
    """usc_malloc'd device memory with numpy round trips (for tests and the plain-C style host)."""

    def __init__(self, handle, nbytes):
        self.h = handle
        self.nbytes = int(nbytes)
        p = C.c_void_p()
        _ck(load().usc_malloc_on(handle._h, C.byref(p), C.c_size_t(max(self.nbytes, 1))))
        self.ptr = p.value

    @classmethod
    def from_numpy(cls, handle, arr):
        arr = np.ascontiguousarray(arr)
        b = cls(handle, arr.nbytes)
        if arr.nbytes:
            _ck(load().usc_memcpy_h2d(handle._h, C.c_void_p(b.ptr), arr.ctypes.data_as(C.c_void_p), C.c_size_t(arr.nbytes)))
            handle.sync()
        return b

    def to_numpy(self, dtype, count=None):
        dtype = np.dtype(dtype)
        n = self.nbytes // dtype.itemsize if count is None else count
        out = np.empty(n, dtype)
        if out.nbytes:
            _ck(load().usc_memcpy_d2h(self.h._h, out.ctypes.data_as(C.c_void_p), C.c_void_p(self.ptr), C.c_size_t(out.nbytes)))
            self.h.sync()
        return out

    def free(self):
        if self.ptr:
            load().usc_free(C.c_void_p(self.ptr))
            self.ptr = 0

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


def default_config(**over):
    cfg = Config()
    load().usc_default_config(C.byref(cfg))
    for k, v in over.items():
        setattr(cfg, k, v)
    return cfg


class Handle:
    """usc_handle: tables + stream for one receiver configuration on one GPU."""

    def __init__(self, cfg=None, device=0, **over):
        L = load()
        self.cfg = cfg if cfg is not None else default_config(**over)
        self._h = C.c_void_p()
        _ck(L.usc_create(C.byref(self.cfg), C.c_int(device), C.byref(self._h)))
        self.n = int(self.cfg.n)

    def close(self):
        if self._h:
            load().usc_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- lifecycle -------------------------------------------------------------------------------
    def set_stream(self, cuda_stream):
        _ck(load().usc_set_stream(self._h, C.c_void_p(int(cuda_stream))))

    def sync(self):
        _ck(load().usc_sync(self._h))

    @property
    def launch_count(self):
        return int(load().usc_launch_count(self._h))

    def geometry(self):
        b, b2, z = C.c_uint32(), C.c_uint32(), C.c_uint32()
        _ck(load().usc_get_geometry(self._h, C.byref(b), C.byref(b2), C.byref(z)))
        return b.value, b2.value, z.value

    def table(self, what):
        n = load().usc_get_table(self._h, what.encode(), None, C.c_size_t(0))
        if n < 0:
            raise UscError(n)
        out = np.empty(n, np.float32)
        r = load().usc_get_table(self._h, what.encode(), out.ctypes.data_as(C.POINTER(C.c_float)), C.c_size_t(n))
        if r < 0:
            raise UscError(r)
        return out

    def buffer(self, arr):
        return DeviceBuffer.from_numpy(self, arr)

    def empty(self, nbytes):
        return DeviceBuffer(self, nbytes)

    # -- batched CMSIS-shaped operators (device pointers) ------------------------------------------
    def i32_to_f32(self, src, dst, count):
        _ck(load().usc_i32_to_f32(self._h, _ptr(src), _ptr(dst), C.c_size_t(count)))

    def arm_mult_f32(self, a, stride_a, b, stride_b, dst, stride_dst, block_size, batch):
        _ck(load().usc_arm_mult_f32_batch(self._h, _ptr(a), C.c_size_t(stride_a), _ptr(b), C.c_size_t(stride_b),
                                          _ptr(dst), C.c_size_t(stride_dst), C.c_uint32(block_size), C.c_uint32(batch)))

    def arm_scale_f32(self, src, scale, dst, block_size, batch):
        _ck(load().usc_arm_scale_f32_batch(self._h, _ptr(src), C.c_float(scale), _ptr(dst), C.c_uint32(block_size),
                                           C.c_uint32(batch)))

    def arm_cmplx_mult_cmplx_f32(self, a, stride_a, b, stride_b, dst, stride_dst, num_samples, batch):
        _ck(load().usc_arm_cmplx_mult_cmplx_f32_batch(self._h, _ptr(a), C.c_size_t(stride_a), _ptr(b),
                                                      C.c_size_t(stride_b), _ptr(dst), C.c_size_t(stride_dst),
                                                      C.c_uint32(num_samples), C.c_uint32(batch)))

    def arm_cmplx_mult_real_f32(self, cplx, stride_c, real, stride_r, dst, stride_dst, num_samples, batch):
        _ck(load().usc_arm_cmplx_mult_real_f32_batch(self._h, _ptr(cplx), C.c_size_t(stride_c), _ptr(real),
                                                     C.c_size_t(stride_r), _ptr(dst), C.c_size_t(stride_dst),
                                                     C.c_uint32(num_samples), C.c_uint32(batch)))

    def arm_cmplx_mag_f32(self, src, stride_src, dst, stride_dst, num_samples, batch):
        _ck(load().usc_arm_cmplx_mag_f32_batch(self._h, _ptr(src), C.c_size_t(stride_src), _ptr(dst),
                                               C.c_size_t(stride_dst), C.c_uint32(num_samples), C.c_uint32(batch)))

    def arm_max_f32(self, src, stride_src, block_size, result, index, batch):
        _ck(load().usc_arm_max_f32_batch(self._h, _ptr(src), C.c_size_t(stride_src), C.c_uint32(block_size),
                                         _ptr(result), _ptr(index), C.c_uint32(batch)))

    def arm_mean_f32(self, src, stride_src, block_size, result, batch):
        _ck(load().usc_arm_mean_f32_batch(self._h, _ptr(src), C.c_size_t(stride_src), C.c_uint32(block_size),
                                          _ptr(result), C.c_uint32(batch)))

    def arm_rfft_fast_f32(self, fft_len, src, dst, ifft_flag, batch):
        _ck(load().usc_arm_rfft_fast_f32_batch(self._h, C.c_uint32(fft_len), _ptr(src), _ptr(dst),
                                               C.c_uint8(1 if ifft_flag else 0), C.c_uint32(batch)))

    def arm_cfft_f32(self, fft_len, data, ifft_flag, batch):
        _ck(load().usc_arm_cfft_f32_batch(self._h, C.c_uint32(fft_len), _ptr(data), C.c_uint8(1 if ifft_flag else 0),
                                          C.c_uint32(batch)))

    def arm_fir_f32(self, coeffs_host, state, src, dst, block_size, batch):
        c = np.ascontiguousarray(coeffs_host, np.float32)
        _ck(load().usc_arm_fir_f32_batch(self._h, c.ctypes.data_as(C.POINTER(C.c_float)), C.c_uint32(c.size),
                                         _ptr(state), _ptr(src), _ptr(dst), C.c_uint32(block_size), C.c_uint32(batch)))

    # -- fused stage-level operators -----------------------------------------------------------------
    def demod_frames(self, pcm, pcm_format, nframes, mag_up=None, idx_up=None, mag_down=None, idx_down=None,
                     bit=None):
        _ck(load().usc_demod_frames(self._h, _ptr(pcm), C.c_uint32(pcm_format), C.c_size_t(nframes), _ptr(mag_up),
                                    _ptr(idx_up), _ptr(mag_down), _ptr(idx_down), _ptr(bit)))

    @staticmethod
    def _hp(x):
        """raw HOST address of an int / numpy array / (pinned) torch tensor"""
        if x is None:
            return C.c_void_p(0)
        if isinstance(x, int):
            return C.c_void_p(x)
        if isinstance(x, np.ndarray):
            return C.c_void_p(x.ctypes.data)
        return C.c_void_p(x.data_ptr())

    def iq_demod_hostbuf(self, pcm_host, pcm_format, nstreams, nframes, stream_stride, mag_up, idx_up, mag_down, idx_down, bit):
        hp = self._hp
        _ck(load().usc_iq_demod_host(self._h, hp(pcm_host), C.c_uint32(pcm_format), C.c_uint32(nstreams), C.c_uint32(nframes),
                                     C.c_size_t(stream_stride), hp(mag_up), hp(idx_up), hp(mag_down), hp(idx_down), hp(bit)))

    def receiver_run_hostbuf(self, pcm_host, pcm_format, nstreams, nframes, stream_stride, uart, uart_cap, results):
        hp = self._hp
        _ck(load().usc_receiver_run_host(self._h, hp(pcm_host), C.c_uint32(pcm_format), C.c_uint32(nstreams), C.c_uint32(nframes),
                                         C.c_size_t(stream_stride), hp(uart), C.c_uint32(uart_cap), hp(results)))

    def host_workspace(self, chunk_frames):
        _ck(load().usc_host_workspace(self._h, C.c_size_t(chunk_frames)))

    def demod_frames_hostbuf(self, pcm_host_ptr, pcm_format, nframes, mag_up, idx_up, mag_down, idx_down, bit):
        """usc_demod_frames_host on raw HOST addresses (ints / numpy arrays / pinned torch tensors)."""
        hp = self._hp
        _ck(load().usc_demod_frames_host(self._h, hp(pcm_host_ptr), C.c_uint32(pcm_format), C.c_size_t(nframes),
                                         hp(mag_up), hp(idx_up), hp(mag_down), hp(idx_down), hp(bit)))

    def receiver_run(self, pcm, pcm_format, nstreams, nframes, stream_stride, uart, uart_cap, results):
        _ck(load().usc_receiver_run(self._h, _ptr(pcm), C.c_uint32(pcm_format), C.c_uint32(nstreams), C.c_uint32(nframes),
                                    C.c_size_t(stream_stride), _ptr(uart), C.c_uint32(uart_cap), _ptr(results)))

    def receiver_run_chunk(self, pcm, pcm_format, nstreams, nframes, stream_stride, carry_frames, state, uart, uart_cap, results):
        _ck(load().usc_receiver_run_chunk(self._h, _ptr(pcm), C.c_uint32(pcm_format), C.c_uint32(nstreams), C.c_uint32(nframes),
                                          C.c_size_t(stream_stride), C.c_uint32(carry_frames), _ptr(state), _ptr(uart),
                                          C.c_uint32(uart_cap), _ptr(results)))

    def sync_search(self, pcm, pcm_format, nstreams, nframes, stream_stride, sync_add, mag, idx):
        _ck(load().usc_sync_search(self._h, _ptr(pcm), C.c_uint32(pcm_format), C.c_uint32(nstreams), C.c_uint32(nframes),
                                   C.c_size_t(stream_stride), C.c_uint32(sync_add), _ptr(mag), _ptr(idx)))

    def receiver_run_host(self, pcm, uart_cap=256):
        """numpy convenience: pcm [nstreams, nframes, n] int32 -> (list of uart byte strings, results array)."""
        pcm = np.ascontiguousarray(pcm, dtype=np.int32)
        S, F = pcm.shape[0], pcm.shape[1]
        d = self.buffer(pcm)
        d_u, d_r = self.empty(S * uart_cap), self.empty(S * rx_result_dtype.itemsize)
        self.receiver_run(d, PCM_I32, S, F, F * self.n, d_u, uart_cap, d_r)
        self.sync()
        res = d_r.to_numpy(rx_result_dtype)
        u = d_u.to_numpy(np.uint8).reshape(S, uart_cap)
        return [bytes(u[s, :min(int(res["nbytes"][s]), uart_cap)]) for s in range(S)], res

    def synth_iq_frames(self, seed, first_frame, nframes, carrier, bw, sideband, phase, amp, noise_sigma, pcm, bits=None):
        _ck(load().usc_synth_iq_frames(self._h, C.c_uint64(seed), C.c_uint64(first_frame), C.c_size_t(nframes), C.c_double(carrier),
                                       C.c_double(bw), C.c_int(sideband), C.c_double(phase), C.c_double(amp), C.c_double(noise_sigma),
                                       _ptr(pcm), _ptr(bits)))

    def synth_streams(self, seed, first_stream, nstreams, nframes, stream_stride, lead_in, msg_bytes, guard, amp, noise_sigma,
                      pcm, offsets=None, messages=None):
        _ck(load().usc_synth_streams(self._h, C.c_uint64(seed), C.c_uint64(first_stream), C.c_uint32(nstreams),
                                     C.c_uint32(nframes), C.c_size_t(stream_stride), C.c_uint32(lead_in), C.c_uint32(msg_bytes),
                                     C.c_uint32(guard), C.c_double(amp), C.c_double(noise_sigma), _ptr(pcm), _ptr(offsets),
                                     _ptr(messages)))

    def synth_frames(self, seed, first_frame, nframes, amp, noise_sigma, pcm, bits=None):
        _ck(load().usc_synth_frames(self._h, C.c_uint64(seed), C.c_uint64(first_frame), C.c_size_t(nframes), C.c_double(amp),
                                    C.c_double(noise_sigma), _ptr(pcm), _ptr(bits)))

    def scan4(self, pcm2n, batch, out):
        _ck(load().usc_scan4(self._h, _ptr(pcm2n), C.c_uint32(batch), _ptr(out)))

    def resample_i16_to_pcm(self, src, n_in, up, down, dst, n_out):
        _ck(load().usc_resample_i16_to_pcm(self._h, _ptr(src), C.c_size_t(n_in), C.c_uint32(up), C.c_uint32(down), _ptr(dst),
                                           C.c_size_t(n_out)))

    def band_magnitudes(self, pcm, pcm_format, nframes, mag):
        _ck(load().usc_band_magnitudes(self._h, _ptr(pcm), C.c_uint32(pcm_format), C.c_size_t(nframes), _ptr(mag)))

    def onoff_detect(self, pcm, pcm_format, nstreams, nframes, cfg=None, strength=None, level=None, chars=None, cap=0,
                     nchars=None, sync_errors=None):
        if cfg is None:
            cfg = OnOffConfig()
            load().usc_onoff_default_config(C.byref(cfg))
        _ck(load().usc_onoff_detect(self._h, _ptr(pcm), C.c_uint32(pcm_format), C.c_uint32(nstreams), C.c_uint32(nframes),
                                    C.byref(cfg), _ptr(strength), _ptr(level), _ptr(chars), C.c_uint32(cap), _ptr(nchars),
                                    _ptr(sync_errors)))

    def fsk_detect(self, pcm, pcm_format, nstreams, nframes, cfg=None, code=None, magnitude=None, frequency=None, chars=None,
                   cap=0, nchars=None, nsof=None, neof=None):
        if cfg is None:
            cfg = FskConfig()
            load().usc_fsk_default_config(C.byref(cfg))
        _ck(load().usc_fsk_detect(self._h, _ptr(pcm), C.c_uint32(pcm_format), C.c_uint32(nstreams), C.c_uint32(nframes),
                                  C.byref(cfg), _ptr(code), _ptr(magnitude), _ptr(frequency), _ptr(chars), C.c_uint32(cap),
                                  _ptr(nchars), _ptr(nsof), _ptr(neof)))

    def spectrum_analyzer(self, pcm, pcm_format, nframes, ac_coupling_hz, mag=None, db=None, peak=None, peak_idx=None):
        _ck(load().usc_spectrum_analyzer(self._h, _ptr(pcm), C.c_uint32(pcm_format), C.c_uint32(nframes),
                                         C.c_float(ac_coupling_hz), _ptr(mag), _ptr(db), _ptr(peak), _ptr(peak_idx)))

    def iq_init(self, carrier_hz, bw_hz, fir_coeffs, window_bins=32):
        c = np.ascontiguousarray(fir_coeffs, np.float32)
        _ck(load().usc_iq_init(self._h, C.c_float(carrier_hz), C.c_float(bw_hz), c.ctypes.data_as(C.POINTER(C.c_float)),
                               C.c_uint32(c.size), C.c_uint32(window_bins)))

    def iq_demod(self, pcm, pcm_format, nstreams, nframes, stream_stride, mag_up=None, idx_up=None, mag_down=None,
                 idx_down=None, bit=None):
        _ck(load().usc_iq_demod(self._h, _ptr(pcm), C.c_uint32(pcm_format), C.c_uint32(nstreams), C.c_uint32(nframes),
                                C.c_size_t(stream_stride), _ptr(mag_up), _ptr(idx_up), _ptr(mag_down), _ptr(idx_down),
                                _ptr(bit)))

    def pipeline(self, frames, mags, updown, batch):
        _ck(load().usc_pipeline(self._h, _ptr(frames), _ptr(mags), C.c_int(updown), C.c_uint32(batch)))

    def dsp(self, fifo, fifo_stride, sync_position, mag_mean, updown, hist, batch):
        _ck(load().usc_dsp(self._h, _ptr(fifo), C.c_size_t(fifo_stride), _ptr(sync_position), _ptr(mag_mean),
                           C.c_int(updown), _ptr(hist), C.c_uint32(batch)))

    def compress_chirp(self, pcm, pcm_format, nframes, use_up, out_frames=None, max_val=None, max_idx=None):
        _ck(load().usc_compress_chirp(self._h, _ptr(pcm), C.c_uint32(pcm_format), C.c_size_t(nframes),
                                      C.c_int(1 if use_up else 0), _ptr(out_frames), _ptr(max_val), _ptr(max_idx)))

    def correlate_os(self, pcm, pcm_format, nstreams, nframes, stream_stride, use_up, out=None, max_val=None, max_idx=None):
        _ck(load().usc_correlate_os(self._h, _ptr(pcm), C.c_uint32(pcm_format), C.c_uint32(nstreams), C.c_uint32(nframes),
                                    C.c_size_t(stream_stride), C.c_int(1 if use_up else 0), _ptr(out), _ptr(max_val), _ptr(max_idx)))

    # -- numpy convenience (host arrays in, host arrays out; used by tests) --------------------------
    def demod_frames_host(self, pcm):
        pcm = np.ascontiguousarray(pcm)
        fmt = PCM_I32 if pcm.dtype == np.int32 else PCM_F32
        if fmt == PCM_F32:
            pcm = pcm.astype(np.float32, copy=False)
        nf = pcm.size // self.n
        d_in = self.buffer(pcm)
        outs = [self.empty(4 * nf) for _ in range(4)]
        d_bit = self.empty(nf)
        self.demod_frames(d_in, fmt, nf, outs[0], outs[1], outs[2], outs[3], d_bit)
        self.sync()
        return (outs[0].to_numpy(np.float32), outs[1].to_numpy(np.uint32), outs[2].to_numpy(np.float32),
                outs[3].to_numpy(np.uint32), d_bit.to_numpy(np.uint8))
