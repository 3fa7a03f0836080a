"""CPU, gloo, world_size 2: the sharding + all-gather + merge logic of nope_b200.dist
(the compute itself needs a GPU; here each rank fabricates its shard's scores from a
shared seeded similarity matrix and the merged result must equal the single-process
ranking, on every rank)."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, n_poses, k, ret):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from nope_b200.dist import ShardedSweep, shard_range
        from oracle.unet_oracle import topk_lowest_index
        g = torch.Generator().manual_seed(123)
        sim = torch.randn(3, n_poses, generator=g)
        sim[:, 4] = sim[:, 2]                      # a tie across (possibly) different shards
        sim[0, n_poses - 1] = sim[0].max() + 1      # best in the last shard
        lo, hi = shard_range(n_poses, rank, world)
        local = sim[:, lo:hi].contiguous()
        kl = min(k, hi - lo)
        if kl > 0:
            li = topk_lowest_index(local, kl)
            tv = torch.gather(local, 1, li)
            ti = li + lo
        else:
            tv = torch.empty(3, 0)
            ti = torch.empty(3, 0, dtype=torch.int64)
        ss = ShardedSweep()
        send, recv, topv_v, topi_v, off_s, pack, per = ss.record(3, k, n_poses, True, torch.device("cpu"))
        ss.fill_record(send, topv_v, topi_v, off_s, local, tv, ti)
        full, topi = ss.gather_merge(send, recv, 3, k, n_poses, per, pack, True)     # ONE all-gather
        ok = torch.equal(full, sim) and torch.equal(topi, topk_lowest_index(sim, k))
        # the record layout is the one the CUDA merge kernel reads (include/nope_b200.h)
        from nope_b200 import _lib
        ok = ok and pack == _lib.load().nope_topk_pack_floats(3, k, per, 1)
        ret[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_poses,k", [(26, 5), (7, 5), (5, 5), (643, 5)])
def test_shard_gather_merge_gloo(n_poses, k):
    world = 2
    port = 29500 + (os.getpid() + n_poses) % 2000
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, port, n_poses, k, ret), nprocs=world, join=True)
    assert dict(ret) == {0: True, 1: True}


def test_shard_range_covers_grid():
    from nope_b200.dist import shard_range
    for n in (1, 5, 26, 642, 2562, 10248):
        for w in (1, 2, 4, 8):
            spans = [shard_range(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            for a, b in zip(spans[:-1], spans[1:]):
                assert a[1] == b[0]


def test_merge_topk_ties():
    from nope_b200.dist import merge_topk
    vals = torch.tensor([[1.0], [3.0], [3.0], [2.0], [float("-inf")]])
    idx = torch.tensor([[9], [7], [4], [1], [-1]])
    v, i = merge_topk(vals, idx, 3)
    assert i.tolist() == [[4, 7, 1]] and v.tolist() == [[3.0, 3.0, 2.0]]


def test_record_length_keeps_int64_blocks_aligned():
    """records are concatenated by the all-gather and read in place by the CUDA merge: every record length must be
    a multiple of 2 floats (we pad to 4) for any batch / k / shard size -- 10 248 poses on 8 ranks (1281 per rank)
    once produced 1297 floats and a misaligned-address fault on every second record"""
    from nope_b200.dist import record_layout
    for B in (1, 2, 3, 8):
        for k in (1, 3, 5):
            for per in (1, 81, 321, 642, 1281, 2562):
                for want_sim in (False, True):
                    off_i, off_s, pack = record_layout(B, k, per, want_sim)
                    assert off_i % 2 == 0 and pack % 4 == 0
                    assert off_i >= B * k and off_s == off_i + 2 * B * k
                    assert pack >= off_s + (B * per if want_sim else 0)
