"""Writes tests/golden/oracle_sha256.json: digests of the ORACLE's answer on the BASELINE configurations at their
stated sizes (SURVEY 8c ii), from workloads generated on the CPU (torch's CPU generator: the same bytes on every
machine).  Per configuration: `sha256` over (corrected bases, read boundaries, status) -- what tests compare the CUDA
path and the oracle with -- and `fa_sha256` / `log_sha256`, the digests of the `.fa` (70 columns, ids read_<n>) and of
the SORTED `.log` those arrays format to (the files the `talc` command line writes; sorted because the reference's
log order under -t N is nondeterministic).

The day a real reference binary exists (tools/pin_reference.sh), its `.fa` / `.log` are diffed against these.
Run from the repository root:  python tools/make_pins.py [--with-config2]"""
import hashlib
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402

from oracle import pyoracle as po  # noqa: E402
from talc_b200 import synth  # noqa: E402

MESSAGES = {1: "No solid kmer could be found.", 2: "Unable to define convenient structure."}

# name -> (config index, scale, reads, junctions, overrides)
PINS = {
    "config1": (1, 1.0, 10000, False, {}),
    "config3_small": (3, 0.004, 250, True, {}),
    "config5": (5, 1.0, 10000, False, {}),
    "k31": (5, 0.3, 1500, False, {"k": 31}),
}
BIG = {
    "config2_batch": (2, 1.0, 131072, False, {}),
    "config3_batch": (3, 1.0, 32768, True, {}),
}


def array_digest(out, off, status) -> str:
    h = hashlib.sha256()
    h.update(np.ascontiguousarray(out, dtype=np.uint8).tobytes())
    h.update(np.ascontiguousarray(off, dtype=np.uint64).tobytes())
    h.update(np.ascontiguousarray(status, dtype=np.uint8).tobytes())
    return h.hexdigest()


def fasta_bytes(out, off, ids) -> bytes:
    """io.cpp:50-75 / SeqAn writeRecords: '>' id, the sequence at 70 columns, an empty sequence is one empty line."""
    raw = np.ascontiguousarray(out, dtype=np.uint8).tobytes()
    parts = []
    for r, name in enumerate(ids):
        s = raw[int(off[r]):int(off[r + 1])]
        parts.append(b">" + name + b"\n")
        if not s:
            parts.append(b"\n")
        for j in range(0, len(s), 70):
            parts.append(s[j:j + 70] + b"\n")
    return b"".join(parts)


def log_lines(status, ids):
    return [b"[Read: " + ids[r] + b" ]: " + MESSAGES[int(s)].encode() + b"\n" for r, s in enumerate(status) if int(s) in MESSAGES]


def make(name, spec, threads):
    ci, scale, n, usej, over = spec
    cfg = synth.baseline_config(ci, scale)
    cfg.n_reads = n
    for k, v in over.items():
        setattr(cfg, k, v)
    w = synth.make_workload(cfg)
    keys, counts = w.keys.numpy().astype(np.uint64), w.counts.numpy().astype(np.int64)
    jk = w.jkeys.numpy().astype(np.uint64) if usej else None
    jc = w.jcounts.numpy().astype(np.int64) if usej else None
    reads, off = w.reads.numpy(), w.read_off.numpy().astype(np.uint64)
    t = po.OracleTable(po.make_params(k=cfg.k)).build_packed(keys, counts, jk, jc)
    out, ooff, st, ctr, secs = t.correct(reads, off, threads=threads)
    ids = [b"read_%d" % r for r in range(n)]
    hin = hashlib.sha256()
    for a in (keys, counts, reads, off) + ((jk, jc) if usej else ()):
        hin.update(np.ascontiguousarray(a).tobytes())
    rec = {"input_sha256": hin.hexdigest(), "config": ci, "scale": scale, "reads": n, "k": cfg.k, "junctions": usej, "bases_in": int(off[-1]),
           "bases_out": int(ooff[-1]), "table_entries_kept": t.size(), "sha256": array_digest(out, ooff, st),
           "fa_sha256": hashlib.sha256(fasta_bytes(out, ooff, ids)).hexdigest(),
           "log_sha256": hashlib.sha256(b"".join(sorted(log_lines(st, ids)))).hexdigest(),
           "failed_reads": int(sum(1 for s in st if int(s) in MESSAGES)),
           "counters": {k: ctr[k] for k in ("lookups_walk", "steps_inner", "steps_border", "gaps", "gaps_bridged", "borders",
                                            "borders_corrected", "cells_nw", "cells_xdrop")}}
    print(name, rec["sha256"][:16], "reads", n, "oracle %.1f s" % secs, flush=True)
    return rec


def main():
    path = os.path.join("tests", "golden", "oracle_sha256.json")
    pins = json.load(open(path)) if os.path.exists(path) else {}
    specs = dict(PINS)
    if "--with-config2" in sys.argv:
        specs.update(BIG)
    threads = os.cpu_count() or 8
    for name, spec in specs.items():
        pins[name] = make(name, spec, threads)
    json.dump(pins, open(path, "w"), indent=1, sort_keys=True)
    print("wrote", path)


if __name__ == "__main__":
    main()
