"""CPU test of the multi-process path (world_size 2, gloo): shard ranges cover the batch, the max-over-ranks
reduction and the optional output gather reassemble what a single process computes. The per-frame work is done by
the oracle here (CPU); on GPUs the same plumbing moves DeviceArrays over NCCL."""
import os
import socket
import sys

import numpy as np
import pytest

torch = pytest.importorskip("torch")
import torch.distributed as dist  # noqa: E402
import torch.multiprocessing as mp  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_frames, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, 'tests'))
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from oflibnumpy_b200 import dist as ofd
    from oracle import flowref as R
    assert ofd.rank_world() == (rank, world)
    start, stop = ofd.shard_range(n_frames, rank, world)
    rng = np.random.default_rng(5)
    a = (rng.random((n_frames, 12, 16, 2)).astype(np.float32) - 0.5) * 6
    b = (rng.random((n_frames, 12, 16, 2)).astype(np.float32) - 0.5) * 6
    local = np.stack([R.combine(R.make(a[i], 't'), R.make(b[i], 't'), 3).vecs for i in range(start, stop)]) \
        if stop > start else np.zeros((0, 12, 16, 2), np.float32)
    mx = ofd.max_over_ranks([1.0 + rank, 10.0 - rank])
    assert mx == [float(world), 10.0]
    full = ofd.gather_frames(local, n_frames, dst=0)
    if rank == 0:
        want = np.stack([R.combine(R.make(a[i], 't'), R.make(b[i], 't'), 3).vecs for i in range(n_frames)])
        np.testing.assert_array_equal(full.numpy(), want)
        open(os.path.join(out_dir, 'ok'), 'w').write('ok')
    else:
        assert full is None
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize('n_frames', [5, 2, 1])
def test_two_rank_sharding_and_gather(tmp_path, n_frames):
    port = _free_port()
    mp.spawn(_worker, args=(2, port, n_frames, str(tmp_path)), nprocs=2, join=True)
    assert os.path.exists(os.path.join(str(tmp_path), 'ok'))
