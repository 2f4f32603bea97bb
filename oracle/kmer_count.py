"""CPU restatement (numpy) of what `jellyfish count --mer K` followed by `jellyfish dump -c` produces for the
reference's README recipe (README.md:33-55 of lbroseus/TALC): NON-canonical counts of every window of K consecutive
A/C/G/T letters (either case) of every read; windows holding any other letter (N, ...) are not k-mers.  Jellyfish2 is
a third-party tool that is absent from this image; its counting rule is public and this is it.

TEST INFRASTRUCTURE ONLY (checker for talc_table_count_reads, row f3).  PARITY UNPINNED: no Jellyfish binary to run.
"""
from __future__ import annotations

import numpy as np

_CODE = np.full(256, 4, dtype=np.uint8)
for _i, _ch in enumerate(b"ACGT"):
    _CODE[_ch] = _i
    _CODE[_ch + 32] = _i  # lower case


def read_sequences(path: str):
    """Sequence lines of a 4-line FASTQ or a 2-line FASTA file."""
    with open(path, "rb") as f:
        lines = f.read().split(b"\n")
    if lines and lines[-1] == b"":
        lines.pop()
    if not lines:
        return []
    period = 4 if lines[0][:1] == b"@" else 2
    n = len(lines) - len(lines) % period
    return [lines[i].rstrip(b"\r") for i in range(1, n, period)]


def count_kmers(seqs, k: int):
    """(sorted distinct packed k-mers [uint64, first base most significant], their counts [int64], occurrences)."""
    packed = []
    for s in seqs:
        c = _CODE[np.frombuffer(s, dtype=np.uint8)]
        n = len(c)
        if n < k:
            continue
        bad = (c > 3).astype(np.int64)
        csum = np.concatenate([[0], np.cumsum(bad)])
        valid = (csum[k:] - csum[:n - k + 1]) == 0
        km = np.zeros(n - k + 1, dtype=np.uint64)
        c64 = (c & 3).astype(np.uint64)
        for j in range(k):
            km = (km << np.uint64(2)) | c64[j:n - k + 1 + j]
        packed.append(km[valid])
    if not packed:
        return np.zeros(0, np.uint64), np.zeros(0, np.int64), 0
    allk = np.concatenate(packed)
    keys, counts = np.unique(allk, return_counts=True)
    return keys, counts.astype(np.int64), int(allk.size)
