"""The C-ABI boundary without a GPU: libusc.so loads, exports every symbol include/usc.h declares,
fails loudly without a device, and its host-side (plain C) table builders agree bit for bit with
the oracle's independent restatement."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import usc
from oracle import pyref as R

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def test_header_symbols_all_exported():
    hdr = open(os.path.join(ROOT, "include", "usc.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = sorted(set(re.findall(r"\b(usc_[a-z0-9_]+)\s*\(", hdr)))
    assert len(declared) >= 30
    L = usc.load()
    for name in declared:
        assert hasattr(L, name), "libusc.so does not export " + name
    assert sorted(usc.SYMBOLS) == declared


def test_default_config_is_the_final_receiver():
    cfg = usc.default_config()
    assert (cfg.n, cfg.fs, cfg.f0, cfg.f1) == (2048, 78125.0, 16000.0, 19000.0)
    assert cfg.sweep_T == np.float32(0.0205) and cfg.snr_threshold == 2.0
    assert cfg.chirp_variant == usc.CHIRP_R and cfg.window == usc.HANN_PERIODIC
    assert C.sizeof(usc.Config) == 48
    assert usc.history_dtype.itemsize == 48


@pytest.mark.skipif(_has_gpu(), reason="only meaningful on a box without a GPU")
def test_no_cpu_fallback_create_fails_without_gpu():
    with pytest.raises(usc.UscError) as e:
        usc.Handle()
    assert e.value.code <= -1000            # a CUDA error surfaced, not a silent host path


def test_error_strings():
    L = usc.load()
    assert L.usc_error_string(0) == b"ok"
    assert b"ARM_MATH_ARGUMENT_ERROR" in L.usc_error_string(-1)


def test_create_argument_errors_mirror_arm_status():
    """arm_rfft_fast_init_f32 returns ARM_MATH_ARGUMENT_ERROR (-1) for unsupported lengths
    (arm_math.h:373-382, 2242-2244); usc_create does the same before touching CUDA."""
    L = usc.load()
    for bad_n in (0, 31, 1000, 65536 * 2):
        cfg = usc.default_config(n=bad_n)
        hnd = C.c_void_p()
        assert L.usc_create(C.byref(cfg), 0, C.byref(hnd)) == usc.USC_ERR_ARGUMENT
    cfg = usc.default_config(chirp_variant=9)
    assert L.usc_create(C.byref(cfg), 0, C.byref(C.c_void_p())) == usc.USC_ERR_ARGUMENT
    assert L.usc_create(None, 0, C.byref(C.c_void_p())) == usc.USC_ERR_ARGUMENT


# ---- host table builders (product C code) vs the oracle ------------------------------------------
def _host(name, restype=None, argtypes=None):
    f = getattr(usc.load(), name)
    f.restype = restype
    if argtypes:
        f.argtypes = argtypes
    return f


def test_host_arm_cos_matches_oracle():
    f = _host("usc_host_arm_cos_f32", C.c_float, [C.c_float])
    xs = np.concatenate([np.linspace(-50, 2500, 4001), np.arange(2048) * np.float32(2 * np.pi / 2048)]).astype(np.float32)
    got = np.array([f(float(x)) for x in xs], np.float32)
    assert np.array_equal(got, R.arm_cos_f32(xs))


@pytest.mark.parametrize("n", [32, 256, 2048, 4096])
@pytest.mark.parametrize("kind", [0, 1])
def test_host_hann_matches_oracle(n, kind):
    f = _host("usc_host_hann")
    w = np.empty(n, np.float32)
    f(w.ctypes.data_as(C.c_void_p), C.c_uint32(n), C.c_uint32(kind))
    assert np.array_equal(w, R.hann_window(n, symmetric=bool(kind)))


@pytest.mark.parametrize("variant,fs,f0,f1,T,phase", [
    ("R", 78125.0, 16000.0, 19000.0, 0.0205, -90.0),
    ("S", 78125.0, 16000.0, 19000.0, 0.0205, -90.0),
    ("S", 100000.0, 16000.0, 19000.0, 0.0205, -90.0),
    ("T", 100000.0, 17000.0, 18000.0, 0.0, float(np.float32(-np.float32(3.14159265358979) / 2.0))),
    ("F", 100000.0, 17000.0, 18000.0, 0.0, 0.0),
])
@pytest.mark.parametrize("up", [0, 1])
def test_host_ref_chirp_matches_oracle(variant, fs, f0, f1, T, phase, up):
    f = _host("usc_host_ref_chirp")
    n = 2048
    v = "RSTF".index(variant)
    out = np.empty(2 * n if variant == "S" else n, np.float32)
    f(C.c_uint32(v), C.c_uint32(n), C.c_float(fs), C.c_float(f0), C.c_float(f1), C.c_float(T), C.c_float(phase),
      C.c_int(up), out.ctypes.data_as(C.c_void_p))
    want = R.generate_ref_chirp(variant, n, fs, f0, f1, T, phase, up)
    assert np.array_equal(out, want)
    assert np.abs(out).max() <= 1.0 + 1e-6 and np.abs(out).max() > 0.99


@pytest.mark.parametrize("n", [16, 32, 1024, 2048, 4096, 65536])
def test_host_twiddles_and_radices_match_oracle(n):
    f = _host("usc_host_twiddles")
    tw = np.empty(2 * n, np.float32)
    f(tw.ctypes.data_as(C.c_void_p), C.c_uint32(n))
    for j in (0, 1, n // 4, n // 2 - 1, n - 1, n // 3):
        re, im = R.twiddle(j, n)
        assert tw[2 * j] == re and tw[2 * j + 1] == im
    g = _host("usc_host_radices", C.c_uint32)
    rad = (C.c_uint32 * 8)()
    cnt = g(C.c_uint32(n), rad)
    assert [int(rad[i]) for i in range(cnt)] == R.fft_radices(n)
    assert int(np.prod(R.fft_radices(n))) == n


def test_device_w32_literals_match_table():
    """The base-kernel twiddles are hex-float literals in usc_arith.cuh; they must equal the
    master-table values (cos, -sin)(2*pi*k/32) both builders produce."""
    src = open(os.path.join(ROOT, "ultrasonic-communication_b200", "csrc", "usc_arith.cuh")).read()
    blocks = re.findall(r"constexpr float t\[16\] = \{(.*?)\};", src, flags=re.S)
    assert len(blocks) == 2
    vals = [[float.fromhex(v.strip().rstrip("f")) for v in b.split(",")] for b in blocks]
    for k in range(16):
        re_, im_ = R.twiddle(k, 32)
        assert np.float32(vals[0][k]) == re_, k
        assert np.float32(vals[1][k]) == im_, k


def test_bandwidth_geometry():
    """receiver/Src/main.c:372-374: bandwidth = (F1-F0)*NN/fs -> 78, bandwidth2 = 156, left zero = 1892."""
    f = _host("usc_host_bandwidth", C.c_uint32)
    assert f(C.c_uint32(2048), C.c_float(78125.0), C.c_float(16000.0), C.c_float(19000.0)) == 78
    rx = R.RefReceiver()
    assert (rx.rx.bandwidth, rx.bandwidth2, rx.idx_left_zero) == (78, 156, 1892)
