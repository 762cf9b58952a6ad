"""Cluster kernel only (32768 / 65536 points), for compute-sanitizer racecheck."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "ultrasonic-communication_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import usc, synth
for n, nf in ((32768, 70), (65536, 40)):
    hl = usc.Handle(usc.default_config(n=n))
    pl, _ = synth.make_frames(nf, n=n, seed_noise=n)
    hl.demod_frames_host(pl); hl.close()
print("cluster workload done")
