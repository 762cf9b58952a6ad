"""The stream generator's CPU twin (oracle/ref_dsp.c: ref_synth_streams; frame format of
generator/ChirpGenerator.ipynb cell 2) — structure, reproducibility, and the loop-back through the
oracle receiver (experiments/EXPERIMENT4.md:29 'Hello World!' style)."""
import numpy as np

from oracle import pyref as R

N = 2048


def test_stream_structure_without_noise():
    lead, mb, guard, amp = 5, 2, 3, 1000.0
    pcm, offs, msgs = R.synth_streams(3, 0, 3, 40, lead, mb, guard, amp, 0.0)
    sym, _ = R.synth_frames(3, 0, 1, amp, 0.0)                    # any frame: table values without noise
    tabs = {}
    f = 0
    while len(tabs) < 2:                                          # find one up and one down frame
        p, b = R.synth_frames(3, f, 1, amp, 0.0)
        tabs[int(b[0])] = p[0]
        f += 1
    pattern = lead + 8 + 8 * mb + guard
    for s in range(3):
        x = pcm[s].reshape(-1)
        off = int(offs[s])
        assert off < N and np.all(x[:off] == 0)
        bits = np.unpackbits(msgs[s])
        kinds = [0] * lead + [1] * 7 + [2] + [1 if b else 2 for b in bits] + [0] * guard
        for k in range((x.size - off) // N):
            seg = x[off + k * N: off + (k + 1) * N]
            kind = kinds[k % pattern]
            want = np.zeros(N, np.int32) if kind == 0 else tabs[1 if kind == 1 else 0]
            assert np.array_equal(seg, want), (s, k)
        assert all(0x20 <= c < 0x7f for c in msgs[s])


def test_streams_are_keyed_by_global_index():
    a = R.synth_streams(9, 10, 4, 8, 2, 1, 2, 2.0e4, 3000.0)
    b = R.synth_streams(9, 12, 2, 8, 2, 1, 2, 2.0e4, 3000.0)
    for x, y in zip(a, b):
        assert np.array_equal(x[2:], y)
    c = R.synth_streams(10, 10, 4, 8, 2, 1, 2, 2.0e4, 3000.0)
    assert not np.array_equal(a[0], c[0])


def test_generated_streams_decode_through_the_oracle_receiver():
    pcm, offs, msgs = R.synth_streams(7, 100, 6, 120, 40, 6, 12, 2.0e4, 2000.0)
    rx = R.RefReceiver()
    for s in range(6):
        out, st = R.receiver_run(rx, pcm[s], cap=64)
        assert out == bytes(msgs[s]) + b"\n", s
        assert st.lock_frame > 40
