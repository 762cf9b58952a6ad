"""Multi-GPU host logic on CPU: world_size-2 gloo processes each demodulate their contiguous stream
shard (the oracle stands in for the kernels here) and the gathered symbol vectors equal the
unsharded result in stream order.  No data-path collective exists; only results are gathered."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_ranges_partition_everything():
    from usc.shard import shard_range
    for n in (0, 1, 7, 4096, 262144, 262147):
        for world in (1, 2, 3, 4, 8):
            blocks = [shard_range(n, world, r) for r in range(world)]
            assert blocks[0][0] == 0 and sum(c for _, c in blocks) == n
            for (s0, c0), (s1, _) in zip(blocks, blocks[1:]):
                assert s0 + c0 == s1
            assert max(c for _, c in blocks) - min(c for _, c in blocks) <= 1
    with pytest.raises(ValueError):
        shard_range(4, 2, 2)


def _worker(rank, world, port, nstreams, frames_per_stream, out_path):
    for p in (ROOT, os.path.join(ROOT, "ultrasonic-communication_b200"), os.path.join(ROOT, "tests")):
        sys.path.insert(0, p)
    import synth
    from oracle import pyref as R
    from usc.shard import gather_results, shard_range
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    pcm, _ = synth.make_frames(nstreams * frames_per_stream, snr_db=0.0)
    pcm = pcm.reshape(nstreams, frames_per_stream, -1)
    start, count = shard_range(nstreams, world, rank)
    rx = R.RefReceiver()
    mu, iu, md, idn = rx.demod_frames(pcm[start:start + count].reshape(-1, 2048))
    local = torch.from_numpy(np.stack([iu.astype(np.int64), idn.astype(np.int64), (~(md > mu)).astype(np.int64)], 1)
                             .reshape(count, frames_per_stream, 3))
    full = gather_results(local, nstreams)
    if rank == 0:
        torch.save(full, out_path)
    dist.barrier()
    dist.destroy_process_group()


def test_world_size_2_gloo_gather_matches_unsharded(tmp_path):
    import synth
    from oracle import pyref as R
    nstreams, fps = 5, 3                       # ragged: 3 + 2 streams
    out = str(tmp_path / "gathered.pt")
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, nstreams, fps, out), nprocs=2, join=True)
    full = torch.load(out).numpy()
    pcm, _ = synth.make_frames(nstreams * fps, snr_db=0.0)
    mu, iu, md, idn = R.RefReceiver().demod_frames(pcm)
    want = np.stack([iu.astype(np.int64), idn.astype(np.int64), (~(md > mu)).astype(np.int64)], 1).reshape(nstreams, fps, 3)
    assert np.array_equal(full, want)
