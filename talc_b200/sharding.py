"""Host-side sharding of a read batch over the GPUs of one box: contiguous blocks of reads,
balanced by bases, one table replica per GPU, no per-read communication; outputs are concatenated back in
input order.  The same rule is used by the `talc` CLI (csrc/host/talc_main.cpp)."""
from __future__ import annotations

import numpy as np


def shard_bounds(offsets: np.ndarray, world: int):
    """cut[g]..cut[g+1] are the reads of rank g; cuts fall where the running base count passes g/world."""
    offsets = np.asarray(offsets, dtype=np.uint64)
    n = len(offsets) - 1
    total = int(offsets[n])
    cut = [0]
    r = 0
    for g in range(1, world):
        want = total // world * g
        while r < n and int(offsets[r]) < want:
            r += 1
        cut.append(r)
    cut.append(n)
    return cut


def take_shard(reads: np.ndarray, offsets: np.ndarray, lo: int, hi: int):
    offsets = np.asarray(offsets, dtype=np.uint64)
    b0, b1 = int(offsets[lo]), int(offsets[hi])
    return reads[b0:b1], (offsets[lo:hi + 1] - offsets[lo]).astype(np.uint64)


def merge_shards(parts):
    """parts: list of (out_bytes, out_offsets, status) in rank order -> one batch in input order."""
    outs, offs, sts = [], [np.zeros(1, dtype=np.uint64)], []
    base = 0
    for out, off, st in parts:
        outs.append(np.asarray(out, dtype=np.uint8))
        off = np.asarray(off, dtype=np.uint64)
        offs.append(off[1:] + np.uint64(base))
        base += int(off[-1])
        sts.append(np.asarray(st, dtype=np.uint8))
    return np.concatenate(outs) if outs else np.zeros(0, np.uint8), np.concatenate(offs), \
        np.concatenate(sts) if sts else np.zeros(0, np.uint8)
