"""Which reads are slow, and why: per-read SM cycles of the correction kernel (TALC_DEBUG_CYCLES) next to the
oracle's algorithmic counters for the same reads (tuning aid; the oracle is only the measuring stick here)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["TALC_DEBUG_CYCLES"] = "1"
os.environ["TALC_DEBUG_CYCLES_FILE"] = "/tmp/talc_kc.bin"
import numpy as np, torch
from talc_b200 import api, synth
from oracle import pyoracle as po

n_reads = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
cfg = synth.baseline_config(2, 0.02)
cfg.n_reads = n_reads
w = synth.make_workload(cfg, device="cuda")
keys = w.keys.cpu().numpy().astype(np.uint64)
counts = w.counts.cpu().numpy()
t = api.Talc(api.default_params(cfg.k))
t.load_packed(keys, counts)
n, total = w.n_reads(), w.total_bases()
d_reads = w.reads.contiguous()
d_off = w.read_off.to(torch.int64).contiguous()
d_out = torch.empty(2 * total + 64 * n + 4096, dtype=torch.uint8, device="cuda")
d_ooff = torch.zeros(n + 1, dtype=torch.int64, device="cuda")
d_st = torch.zeros(n, dtype=torch.uint8, device="cuda")
t.correct_device(d_reads, d_off, total, d_out, d_ooff, d_st)
kc = np.fromfile("/tmp/talc_kc.bin", dtype=np.uint32).astype(np.float64) / 1024.0
reads = w.reads.cpu().numpy()
off = w.read_off.cpu().numpy().astype(np.uint64)
ot = po.OracleTable(po.make_params(k=cfg.k)).build_packed(keys, counts.astype(np.int64), None, None)
KEYS = ("steps_inner", "steps_border", "frontier_sum", "cells_nw", "cells_lcs", "cells_ovl", "cells_xdrop", "gaps",
        "gap_attempts", "ev_gardening", "ev_bridge", "ev_edge")


def counters(idx):
    parts = [reads[int(off[r]):int(off[r + 1])] for r in idx]
    so = np.zeros(len(idx) + 1, dtype=np.uint64)
    so[1:] = np.cumsum([len(x) for x in parts])
    return ot.correct(np.concatenate(parts), so, threads=1)[3]


order = np.argsort(-kc)
print("mean Mcycles %.2f" % kc.mean())
for r in order[:8]:
    c = counters([int(r)])
    print("read %d len %d  %.1f Mcycles  %s" % (r, off[r + 1] - off[r], kc[r], {k: c[k] for k in KEYS if c.get(k)}))
mid = order[n // 2 - 50:n // 2 + 50]
c = counters([int(r) for r in mid])
print("median-100: %.1f Mcycles  %s" % (kc[mid].mean(), {k: c[k] / 100 for k in KEYS if c.get(k)}))
# least-squares attribution of cycles to counters over a random sample
rng = np.random.default_rng(0)
samp = rng.choice(n, 400, replace=False)
X = []
for r in samp:
    c = counters([int(r)])
    X.append([c.get(k, 0) for k in KEYS] + [1.0])
X = np.array(X, dtype=np.float64)
y = kc[samp]
coef, *_ = np.linalg.lstsq(X, y, rcond=None)
for k, v, m in zip(list(KEYS) + ["const"], coef, X.mean(0)):
    print("  %-14s coef %.3e Mcycles/unit   mean contribution %.2f Mcycles" % (k, v, v * m))
