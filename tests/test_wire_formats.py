"""The reference's CSV wire formats (include/usc_wire.h, host/usc_wire.c): the writers
reproduce the captured files' text, the readers take the captures back, and the history tables of
print_history() round-trip."""
import ctypes as C
import json
import os
import re

import numpy as np
import pytest

import usc

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
LIB = os.path.join(ROOT, "ultrasonic-communication_b200", "libusc_wire.so")


@pytest.fixture(scope="module")
def wire():
    if not os.path.exists(LIB):
        from importlib import import_module
        import_module("build").build()
    L = C.CDLL(LIB)
    libc = C.CDLL(None)
    libc.fopen.restype = C.c_void_p
    libc.fopen.argtypes = [C.c_char_p, C.c_char_p]
    libc.fclose.argtypes = [C.c_void_p]
    return L, libc


def _fp(a):
    return a.ctypes.data_as(C.c_void_p)


def test_header_symbols_exported(wire):
    L, _ = wire
    hdr = open(os.path.join(ROOT, "include", "usc_wire.h")).read()
    names = set(re.findall(r"\b(usc_wire_\w+)\s*\(", hdr))
    assert len(names) == 10
    for n in names:
        getattr(L, n)


def test_writers_reproduce_the_captured_files(wire, device_triples, tmp_path):
    L, _ = wire
    ex = json.load(open(os.path.join(HERE, "golden", "wire_excerpt.json")))
    i = [str(n).endswith(ex["capture"]) for n in device_triples["names"]].index(True)
    raw = device_triples["raw"][i].astype(np.int32)
    flt = device_triples["flt"][i].astype(np.float32)
    mag = device_triples["fft_mag"][i].astype(np.float32)
    db = (10.0 * np.log10(device_triples["fft_mag"][i])).astype(np.float32)
    fs = 1000.0 * float(re.search(r"_([0-9.]+)\(kHz\)", ex["capture"]).group(1))      # the capture's name carries fs
    paths = {e: str(tmp_path / ("x." + e)).encode() for e in ("raw", "flt", "fft")}
    assert L.usc_wire_write_raw(paths["raw"], _fp(raw), C.c_uint32(2048)) == 2048
    assert L.usc_wire_write_flt(paths["flt"], _fp(flt), C.c_uint32(2048)) == 2048
    assert L.usc_wire_write_fft(paths["fft"], C.c_float(fs), C.c_uint32(2048), _fp(mag), _fp(db)) == 1024
    for e in ("raw", "flt", "fft"):
        lines = open(paths[e].decode()).read().splitlines()
        assert lines[0] == ex[e]["header"] and len(lines) == ex[e]["nlines"]
        for r, text in ex[e]["rows"].items():
            if e == "fft":
                # frequency and magnitude columns are exact; dB is recomputed here from the 6-decimal magnitude
                assert lines[int(r)].split(",")[:2] == text.split(",")[:2], (e, r)
                assert abs(float(lines[int(r)].split(",")[2]) - float(text.split(",")[2])) < 1e-4
            else:
                assert lines[int(r)] == text, (e, r)
    # readers
    r2, f2 = np.zeros(2048, np.int32), np.zeros(2048, np.float32)
    fr, mg, d2 = (np.zeros(1024, np.float32) for _ in range(3))
    assert L.usc_wire_read_raw(paths["raw"], C.c_uint32(2048), _fp(r2)) == 2048 and np.array_equal(r2, raw)
    assert L.usc_wire_read_flt(paths["flt"], C.c_uint32(2048), _fp(f2)) == 2048 and np.array_equal(f2, flt)
    assert L.usc_wire_read_fft(paths["fft"], C.c_uint32(1024), _fp(fr), _fp(mg), _fp(d2)) == 1024
    assert np.array_equal(mg, mag) and np.allclose(fr, device_triples["fft_freq"][i], atol=0.051)
    assert L.usc_wire_read_raw(b"/nonexistent/file.raw", C.c_uint32(4), _fp(r2)) == -2


def test_dump_round_trip(wire, tmp_path):
    L, libc = wire
    rng = np.random.default_rng(5)
    n = 256
    pcm = (rng.integers(-2 ** 23, 2 ** 23, n) * 256).astype(np.int32)
    win = (pcm.astype(np.float32) * rng.random(n).astype(np.float32))
    mag = (rng.random(n // 2) * 1e4).astype(np.float32)
    db = (10 * np.log10(mag)).astype(np.float32)
    path = str(tmp_path / "dump.txt").encode()
    f = libc.fopen(path, b"w")
    assert L.usc_wire_write_dump(C.c_void_p(f), b"M1", C.c_float(78125.0), C.c_uint32(n), _fp(mag), _fp(db), _fp(pcm), _fp(win)) == n // 2
    libc.fclose(C.c_void_p(f))
    text = open(path.decode()).read().splitlines()
    assert text[1] == "MEMS mic: M1" and text[3] == "Frequency(Hz),Magnitude,Magnitude(dB)"
    assert text[2].startswith("Frequency at max magnitude: %.1f, Max magnitude: " % (np.argmax(mag) * 78125.0 / n))
    assert "EORAW" in text and text[-1] == "EOFLT"
    f = libc.fopen(path, b"r")
    mic = C.create_string_buffer(16)
    fr, mg, d2 = (np.zeros(n // 2, np.float32) for _ in range(3))
    p2, w2 = np.zeros(n, np.int32), np.zeros(n, np.float32)
    assert L.usc_wire_read_dump(C.c_void_p(f), C.c_uint32(n), mic, C.c_size_t(16), _fp(fr), _fp(mg), _fp(d2), _fp(p2), _fp(w2)) == n // 2
    libc.fclose(C.c_void_p(f))
    assert mic.value == b"M1" and np.array_equal(p2, pcm)
    assert np.allclose(mg, mag, rtol=0, atol=1e-3) and np.allclose(w2, win, rtol=1e-7, atol=1e-2)
    # a truncated dump is a format error
    open(path.decode(), "w").write("\n".join(text[:20]) + "\n")
    f = libc.fopen(path, b"r")
    assert L.usc_wire_read_dump(C.c_void_p(f), C.c_uint32(n), None, C.c_size_t(0), None, None, None, None, None) == -3
    libc.fclose(C.c_void_p(f))


def test_history_tables(wire, tmp_path):
    L, libc = wire
    H = np.zeros(3, dtype=usc.history_dtype)
    H["rank"] = [ord("1"), ord("-"), ord("L")]
    H["snr"] = [12.34, -3.21, 100.0]
    H["max_freq"] = [16500, 17000, 18999]
    H["max_freq_left"] = [0, 1, 2]
    H["max_freq_right"] = [16500, 17000, 18999]
    H["mag_max"] = [1.5e9, 2.25e8, 3.0e7]
    H["mag_max_right"] = H["mag_max"]
    H["mag_mean"] = [1.0e8, 1.0e8, 2.0e8]
    for detail in (0, 1):
        path = str(tmp_path / ("h%d.txt" % detail)).encode()
        f = libc.fopen(path, b"w")
        assert L.usc_wire_write_history(C.c_void_p(f), C.c_int(detail), C.c_uint32(1), C.c_uint32(2), _fp(H), C.c_uint32(3)) == 3
        libc.fclose(C.c_void_p(f))
        text = open(path.decode()).read().splitlines()
        if detail == 0:
            assert text[0] == "G => S" and text[1] == "1,  12.3" and text[2] == "-,  -3.2" and text[3] == "L, 100.0"
        else:
            assert text[1] == "state: SYNCHRONIZING => SYNCHRONIZED"
            assert text[2] == "r,  freq,freq_l,freq_r, t_s, t_f,      max,    max_l,    max_r, mag_mean,   snr"
            assert text[3] == "1, 16500,     0, 16500,   0,   0, 1.50e+09, 0.00e+00, 1.50e+09, 1.00e+08,  12.3"
        f = libc.fopen(path, b"r")
        G = np.zeros(4, dtype=usc.history_dtype)
        ps, st = C.c_uint32(9), C.c_uint32(9)
        assert L.usc_wire_read_history(C.c_void_p(f), C.c_int(detail), C.byref(ps), C.byref(st), _fp(G), C.c_uint32(4)) == 3
        libc.fclose(C.c_void_p(f))
        assert (ps.value, st.value) == (1, 2)
        assert list(G["rank"][:3]) == list(H["rank"])
        assert np.allclose(G["snr"][:3], np.round(H["snr"], 1), atol=1e-6)
        if detail:
            assert list(G["max_freq"][:3]) == list(H["max_freq"]) and np.allclose(G["mag_max"][:3], H["mag_max"], rtol=5e-3)
