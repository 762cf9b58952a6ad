"""The tuned CPU form of the receiver chain (oracle/ref_fast.c, the CPU arm of bench.py) against the plain
restatement (oracle/ref_dsp.c): bit-identical floats and indices — same operations in the same order, one frame
per SIMD lane.  CPU only."""
import numpy as np
import pytest

import synth
from oracle import pyref as R


@pytest.fixture(scope="module")
def rx():
    return R.RefReceiver()


@pytest.mark.parametrize("nframes", [1, 2, 15, 16, 17, 33, 100])
@pytest.mark.parametrize("dtype", [np.int32, np.float32])
def test_fast_form_equals_plain_form(rx, nframes, dtype):
    pcm, _ = synth.make_frames(nframes, seed_noise=40 + nframes, dtype=dtype)
    want = rx.demod_frames(pcm, nthreads=2)
    got = rx.demod_frames_fast(pcm, nthreads=3)
    for w, g in zip(want, got):
        assert np.array_equal(w.view(np.uint32), g.view(np.uint32))


def test_fast_form_edge_inputs(rx):
    rng = np.random.default_rng(5)
    frames = np.zeros((6, 2048), np.int32)
    frames[1] = 2 ** 31 - 1
    frames[2] = -2 ** 31
    frames[3] = rng.integers(-2 ** 31, 2 ** 31 - 1, 2048, dtype=np.int64).astype(np.int32)
    frames[4, ::2] = 1 << 20                     # ties across bins
    frames[5, 7] = 12345
    want = rx.demod_frames(frames)
    got = rx.demod_frames_fast(frames)
    for w, g in zip(want, got):
        assert np.array_equal(w.view(np.uint32), g.view(np.uint32))


def test_simd_width_reported():
    assert R.lib().ref_fast_simd_width() in (0, 8, 16)
