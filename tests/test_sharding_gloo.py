"""world_size-2 test of the N>1 host logic on the CPU (gloo): rank 0 owns the k-mer table arrays and broadcasts
them once, reads are sharded by bases, every rank corrects its shard with no further communication, and the
shards concatenate back to exactly the single-rank answer.  The per-rank worker is the host-emulated device code
(the build container has no GPU); the GPU version of the same protocol is bench.py --gpus N."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT

GOLD = os.path.join(os.path.dirname(__file__), "golden", "tiny_case.npz")


def _worker(rank, world, port, outdir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests", "hostemu"))
    import pyemu
    from oracle import pyoracle as po
    from talc_b200 import sharding
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = np.load(GOLD)
    k = int(g["c1_k"][0])
    # the table lives on rank 0 only; one broadcast replicates it
    if rank == 0:
        keys = torch.from_numpy(g["c1_keys"].astype(np.int64))
        counts = torch.from_numpy(g["c1_counts"].astype(np.int64))
        meta = torch.tensor([keys.numel()], dtype=torch.int64)
    else:
        meta = torch.zeros(1, dtype=torch.int64)
    dist.broadcast(meta, 0)
    if rank != 0:
        keys = torch.zeros(int(meta[0]), dtype=torch.int64)
        counts = torch.zeros(int(meta[0]), dtype=torch.int64)
    dist.broadcast(keys, 0)
    dist.broadcast(counts, 0)
    reads, off = g["c1_reads"], g["c1_off"]
    cut = sharding.shard_bounds(off, world)
    sub, so = sharding.take_shard(reads, off, cut[rank], cut[rank + 1])
    et = pyemu.EmuTable(pyemu.params_from(po.make_params(k=k)), keys.numpy().astype(np.uint64), counts.numpy())
    out, ooff, st, _ = et.correct(sub, so, arena_bytes=1 << 20, wide=True)
    np.savez(os.path.join(outdir, "rank%d.npz" % rank), out=out, off=ooff, st=st, cut=np.array(cut))
    dist.barrier()
    dist.destroy_process_group()


def test_two_ranks_reproduce_the_single_rank_answer(tmp_path):
    from talc_b200 import sharding
    world = 2
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    g = np.load(GOLD)
    parts = []
    for r in range(world):
        d = np.load(tmp_path / ("rank%d.npz" % r))
        parts.append((d["out"], d["off"], d["st"]))
        cut = d["cut"]
    assert cut[0] == 0 and cut[-1] == len(g["c1_off"]) - 1 and 0 < cut[1] < cut[-1]
    out, off, st = sharding.merge_shards(parts)
    assert np.array_equal(st, g["c1_status"])
    assert np.array_equal(off, g["c1_ooff"]) and np.array_equal(out, g["c1_out"])


def test_shard_bounds_balance_by_bases():
    from talc_b200 import sharding
    off = np.concatenate([[0], np.cumsum([100, 100, 100, 5000, 100, 100, 4000, 100])]).astype(np.uint64)
    for w in (1, 2, 3, 4, 8):
        cut = sharding.shard_bounds(off, w)
        assert cut[0] == 0 and cut[-1] == 8 and all(a <= b for a, b in zip(cut, cut[1:])) and len(cut) == w + 1
