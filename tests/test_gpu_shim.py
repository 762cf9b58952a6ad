"""The CMSIS exact-name host shim (libusc_cmsis.so): a reference-style C caller written against the
arm_math.h names links and runs on the GPU library; its results equal the oracle bit for bit."""
import os
import subprocess

import numpy as np
import pytest

import synth
from oracle import pyref as R

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "ultrasonic-communication_b200")
N = 2048


def test_reference_style_caller_links_and_matches_oracle(tmp_path):
    exe = str(tmp_path / "shim_chain")
    subprocess.check_call(["gcc", "-O2", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "c", "shim_chain.c"),
                           "-o", exe, "-L", PKG, "-lusc_cmsis", "-lusc", "-Wl,-rpath," + PKG])
    rx = R.RefReceiver()
    pcm, _ = synth.make_frames(1, snr_db=0.0, dtype=np.float32)
    frame = pcm[0]
    chirp, hann = rx.table("up_chirp"), rx.table("hann")
    rfft = R.Rfft(N)
    H = rfft(R.arm_mult_f32(R.generate_ref_chirp("R", N, 78125.0, 16000.0, 19000.0, 0.0205, -90.0, 0), hann))
    for name, arr in (("frame", frame), ("chirp", chirp), ("hann", hann), ("H", H)):
        arr.astype(np.float32).tofile(str(tmp_path / (name + ".f32")))
    out = subprocess.check_output([exe] + [str(tmp_path / (n + ".f32")) for n in ("frame", "chirp", "hann", "H")], text=True)
    lines = dict((ln.split()[0], ln.split()[1:]) for ln in out.strip().splitlines())
    assert lines["status"] == ["0"]
    # receiver chain
    mags = R.arm_cmplx_mag_f32(rfft(R.arm_mult_f32(R.arm_mult_f32(frame, chirp), hann)))
    peak, bin_ = R.arm_max_f32(mags[:156])
    assert np.float32(float.fromhex(lines["receiver"][0])) == peak and int(lines["receiver"][1]) == bin_
    assert np.float32(float.fromhex(lines["receiver"][2])) == R.arm_mean_f32(mags[:8])
    # compression chain
    comp = rfft(R.arm_cmplx_mult_cmplx_f32(rfft(R.arm_mult_f32(frame, hann)), H), inverse=True)
    peak, bin_ = R.arm_max_f32(comp)
    assert np.float32(float.fromhex(lines["compress"][0])) == peak and int(lines["compress"][1]) == bin_
    # complex FFT
    z = np.zeros(2 * N, np.float32)
    z[0::2] = frame
    peak, bin_ = R.arm_max_f32(R.arm_cmplx_mag_f32(R.Cfft(N)(z)))
    assert np.float32(float.fromhex(lines["cfft"][0])) == peak and int(lines["cfft"][1]) == bin_


def test_analyser_host_reproduces_a_device_capture(device_triples, tmp_path):
    """analyser_host (plain C over the C-ABI + the CSV wire formats): a captured .raw in, the .fft / .flt
    files and the firmware's UART dump out; the .fft magnitudes match the Cortex-M4's own .fft (<= 1e-4),
    the .flt matches the device's .flt text to the printed precision."""
    import subprocess
    pkg = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "ultrasonic-communication_b200")
    exe = os.path.join(pkg, "host", "analyser_host")
    assert os.path.exists(exe), "run python __graft_entry__.py first"
    i = [k for k, ok in enumerate(device_triples["consistent"]) if ok][0]
    name = str(device_triples["names"][i])
    fs = 1000.0 * float(name.split("(kHz)")[0].split("_")[-1])
    raw = device_triples["raw"][i].astype(np.int64)
    rawf = tmp_path / "cap.raw"
    rawf.write_text("Index,Amplitude\n" + "".join("%d,%d\n" % (k, v) for k, v in enumerate(raw)))
    out = tmp_path / "out"
    r = subprocess.run([exe, str(rawf), str(fs), str(out), "M2A"], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr
    fft = np.loadtxt(str(out) + ".fft", delimiter=",", skiprows=1)
    flt = np.loadtxt(str(out) + ".flt", delimiter=",", skiprows=1)[:, 1]
    dev = device_triples["fft_mag"][i]
    sel = device_triples["fft_freq"][i] >= 1000.0
    big = sel & (dev >= 0.01 * dev[sel].max())
    assert (np.abs(fft[:, 1][big] - dev[big]) / dev[big]).max() < 1e-4
    assert np.all(fft[:, 1][~sel] == 1.0)                                   # AC coupling
    assert np.abs(flt - device_triples["flt"][i]).max() <= 1e-6 * max(1.0, np.abs(flt).max())
    lines = r.stdout.splitlines()
    assert lines[1] == "MEMS mic: M2A" and lines[3] == "Frequency(Hz),Magnitude,Magnitude(dB)" and lines[-1] == "EOFLT"
