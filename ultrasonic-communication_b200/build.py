"""Build libusc.so in-tree (sm_100a only).  Usage: python build.py [--force] [--verbose]

nvcc cross-compiles without a GPU.  Objects go to csrc/_build/ (git-ignored); the shared library
lands next to this file so it travels to the GPU box with the repo snapshot.
"""
import concurrent.futures as cf
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(CSRC, "_build")
LIB = os.path.join(HERE, "libusc.so")

NVCC = os.environ.get("NVCC") or shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-fmad=false", "-Wno-deprecated-gpu-targets",
              "-Xcompiler", "-fPIC,-O2,-ffp-contract=off", "-Xptxas", "-v",
              "--expt-relaxed-constexpr"]
CC_FLAGS = ["-O2", "-ffp-contract=off", "-fPIC", "-std=gnu11", "-Wall"]

CU = ["usc_api.cu", "k_demod.cu", "k_receiver.cu", "k_sync.cu", "k_iq.cu", "k_synth.cu", "k_long.cu", "k_legacy.cu", "k_compress.cu", "k_correlate.cu", "k_fft_generic.cu", "k_fft_warp.cu", "k_elementwise.cu"]
C = ["usc_tables.c"]


def _stale(out, deps):
    if not os.path.exists(out):
        return True
    t = os.path.getmtime(out)
    return any(os.path.getmtime(d) > t for d in deps)


def _run(cmd, log):
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    with open(log, "w") as f:
        f.write(" ".join(cmd) + "\n" + r.stdout)
    if r.returncode != 0:
        raise RuntimeError("build failed: %s\n%s" % (" ".join(cmd), r.stdout[-4000:]))
    return r.stdout


def build(force=False, verbose=False):
    os.makedirs(OBJ, exist_ok=True)
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".h", ".cuh"))]
    headers.append(os.path.join(os.path.dirname(HERE), "include", "usc.h"))
    headers.append(os.path.abspath(__file__))
    jobs, objs = [], []
    for src in CU:
        o = os.path.join(OBJ, src + ".o")
        objs.append(o)
        if force or _stale(o, [os.path.join(CSRC, src)] + headers):
            jobs.append(([NVCC] + NVCC_FLAGS + ["-c", os.path.join(CSRC, src), "-o", o], o + ".log"))
    for src in C:
        o = os.path.join(OBJ, src + ".o")
        objs.append(o)
        if force or _stale(o, [os.path.join(CSRC, src)] + headers):
            jobs.append((["gcc"] + CC_FLAGS + ["-c", os.path.join(CSRC, src), "-o", o], o + ".log"))
    if jobs:
        with cf.ThreadPoolExecutor(max_workers=min(8, len(jobs))) as ex:
            for out in ex.map(lambda j: _run(*j), jobs):
                if verbose:
                    print(out)
    if jobs or force or _stale(LIB, objs):
        _run([NVCC, "-Wno-deprecated-gpu-targets", "-shared", "-o", LIB] + objs + ["-Xcompiler", "-fPIC", "-cudart", "static", "-lm"],
             os.path.join(OBJ, "link.log"))
    # CMSIS exact-name host-pointer shim (tests / drop-in demonstration), plain C over the C-ABI
    shim_src = os.path.join(CSRC, "usc_cmsis_shim.c")
    shim_lib = os.path.join(HERE, "libusc_cmsis.so")
    if force or _stale(shim_lib, [shim_src, LIB] + headers):
        _run(["gcc"] + CC_FLAGS + ["-shared", shim_src, "-o", shim_lib, "-L", HERE, "-lusc", "-Wl,-rpath," + HERE,
                                  "-Wl,-rpath,$ORIGIN"], os.path.join(OBJ, "shim.log"))
    # the reference's CSV wire formats: plain host C, no CUDA (include/usc_wire.h)
    wire_src = os.path.join(HERE, "host", "usc_wire.c")
    wire_lib = os.path.join(HERE, "libusc_wire.so")
    inc = os.path.join(os.path.dirname(HERE), "include")
    tx_src = os.path.join(HERE, "host", "usc_tx.c")                 # transmitter symbols, framing, WAV I/O (include/usc_tx.h)
    if force or _stale(wire_lib, [wire_src, tx_src, os.path.join(inc, "usc_wire.h"), os.path.join(inc, "usc_tx.h"),
                                  os.path.join(inc, "usc.h")]):
        _run(["gcc"] + CC_FLAGS + ["-Wextra", "-shared", "-I", inc, wire_src, tx_src, "-o", wire_lib, "-lm"],
             os.path.join(OBJ, "wire.log"))
    # the plain-C host driver links against the C-ABI only (no CUDA headers): proves the boundary
    host_src = os.path.join(HERE, "host", "receiver_host.c")
    host_bin = os.path.join(HERE, "host", "receiver_host")
    if force or _stale(host_bin, [host_src, LIB, os.path.join(os.path.dirname(HERE), "include", "usc.h")]):
        _run(["gcc", "-O2", "-std=gnu11", "-Wall", "-I", os.path.join(os.path.dirname(HERE), "include"), host_src,
              "-o", host_bin, "-L", HERE, "-lusc", "-Wl,-rpath," + HERE, "-Wl,-rpath,$ORIGIN/.."],
             os.path.join(OBJ, "host.log"))
    ana_src = os.path.join(HERE, "host", "analyser_host.c")
    ana_bin = os.path.join(HERE, "host", "analyser_host")
    if force or _stale(ana_bin, [ana_src, LIB, wire_lib, os.path.join(inc, "usc.h"), os.path.join(inc, "usc_wire.h")]):
        _run(["gcc", "-O2", "-std=gnu11", "-Wall", "-I", inc, ana_src, "-o", ana_bin, "-L", HERE, "-lusc", "-lusc_wire",
              "-Wl,-rpath," + HERE, "-Wl,-rpath,$ORIGIN/.."], os.path.join(OBJ, "host_analyser.log"))
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
