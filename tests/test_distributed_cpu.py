"""Host-side multi-GPU logic under a world-size-2 gloo group on CPU: path-range sharding, identical global path
counters on every rank and the single all-reduce of the fp64 moments.  No kernel runs here."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from common import sm  # noqa: F401  (puts the repo on sys.path)
from sde_mc_b200 import _engine as E


def test_shard_is_a_balanced_contiguous_partition():
    for n in (0, 1, 7, 100, 10 ** 9 + 3):
        for size in (1, 2, 3, 8):
            spans = [E.shard(n, r, size) for r in range(size)]
            assert spans[0][0] == 0
            for (o1, c1), (o2, _) in zip(spans, spans[1:]):
                assert o1 + c1 == o2
            assert spans[-1][0] + spans[-1][1] == n
            counts = [c for _, c in spans]
            assert max(counts) - min(counts) <= 1


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, size, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=size)
    try:
        assert E.world() == (rank, size)
        off, cnt = E.shard(1001, rank, size)
        m = E.Moments("cpu")
        # what a rank's kernel would have accumulated for its path range [off, off+cnt): value = path id
        ids = torch.arange(off, off + cnt, dtype=torch.float64)
        m.buf[0], m.buf[1], m.buf[5] = ids.sum(), (ids * ids).sum(), float(cnt)
        got = m.all_reduce().read()
        out[rank] = (got["sum"], got["sumsq"], got["n"])
    finally:
        dist.destroy_process_group()


def test_two_rank_moment_allreduce_matches_single_process():
    port = _free_port()
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
        ids = torch.arange(0, 1001, dtype=torch.float64)
        want = (float(ids.sum()), float((ids * ids).sum()), 1001.0)
        assert out[0] == want and out[1] == want
    mean, se = E.mean_and_stderr(want[0], want[1], 1001)
    assert abs(mean - 500.0) < 1e-12 and abs(se - float(ids.std()) / 1001 ** 0.5) < 1e-9
