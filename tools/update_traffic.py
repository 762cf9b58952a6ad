#!/usr/bin/env python
"""profiles/k1_traffic.json from an `ncu --set full` summary of the bench kernel (tools/ncu_summary.py output).

bench.py reports roofline.traffic from this file only while the sha1 of the kernel's sources equals the one recorded
here, so a stale figure can never be printed.   python tools/update_traffic.py profiles/r2_k1_dual_ncu_full_summary.json
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

UNITS = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def to_bytes(txt):
    v, u = txt.split()
    return float(v.replace(",", "")) * UNITS[u]


def main():
    src = sys.argv[1]
    d = json.load(open(src))
    rd, wr = to_bytes(d["dram__bytes_read.sum"]), to_bytes(d["dram__bytes_write.sum"])
    out = {"kernel": d["kernel"], "dram_bytes_per_launch": rd + wr, "dram_bytes_read": rd, "dram_bytes_write": wr,
           "source": "%s (ncu --set full, 1 launch, %d frames)" % (os.path.relpath(os.path.abspath(src), ROOT), bench.NFRAMES),
           "algorithmic_bytes_per_launch": bench.ALGO_BYTES_PER_SYMBOL * bench.NFRAMES,
           "kernel_source_sha1": bench.kernel_source_hash()}
    json.dump(out, open(os.path.join(ROOT, "profiles", "k1_traffic.json"), "w"), indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
