"""The scoring primitives in isolation (test_align_kernel: one warp per pair, the same device functions the per-read
program calls) on read-like pairs, for ncu pipe-utilisation captures (integer pipe for stage 4).
Launch order: op 5 nw_lcs_fused, op 2 overlap_score_stripes, op 4 X-drop with X = 8, 40, 100 (register bands S = 1, 2, 4)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from talc_b200 import api

rng = np.random.default_rng(5)
def mutate(s, rate):
    out = bytearray()
    for ch in s:
        u = rng.random()
        if u < rate * 0.4: out.append(int(rng.choice(list(b"ACGT"))))
        elif u < rate * 0.7: out.append(ch); out.append(int(rng.choice(list(b"ACGT"))))
        elif u < rate: continue
        else: out.append(ch)
    return bytes(out) or b"A"
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
L = int(sys.argv[2]) if len(sys.argv) > 2 else 300
a, b = [], []
for i in range(n):
    s = bytes(rng.choice(list(b"ACGT"), L).tolist())
    a.append(mutate(s, 0.10)); b.append(s)
t = api.Talc(api.default_params(21))
t.load_packed(np.array([1, 2, 3], dtype=np.uint64), np.array([5, 5, 5], dtype=np.int64))
cells = sum(len(x) * len(y) for x, y in zip(a, b))
for name, op, aux in (("nw_lcs_fused", 5, 0), ("overlap_score_stripes", 2, 0), ("xdrop X=8", 4, 8), ("xdrop X=40", 4, 40), ("xdrop X=100", 4, 100)):
    t0 = time.time(); t.test_align(op, a, b, aux=aux); dt = time.time() - t0
    print("%-24s %d pairs of ~%d: %.1f ms host wall (%.1f G reference cells/s incl. copies)" % (name, n, L, dt * 1e3, cells / dt / 1e9))
