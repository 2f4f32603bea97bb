"""Small fixed workload for ncu captures: builds the table, runs the correction hot path once or twice."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from talc_b200 import api, synth

n_reads = int(sys.argv[1]) if len(sys.argv) > 1 else 4000
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 1
cfg = synth.baseline_config(int(os.environ.get("TALC_PROFILE_CONFIG", "2")), 0.02)
cfg.n_reads = n_reads
w = synth.make_workload(cfg, device="cuda")
t = api.Talc(api.default_params(cfg.k))
use_j = os.environ.get("TALC_PROFILE_CONFIG", "2") == "3"  # config 3 = config 2 + the junction k-mer dump
t.load_packed(w.keys.cpu().numpy().astype(np.uint64), w.counts.cpu().numpy(),
              w.jkeys.cpu().numpy().astype(np.uint64) if use_j else None, w.jcounts.cpu().numpy() if use_j else None)
n, total = w.n_reads(), w.total_bases()
d_reads = w.reads.contiguous()
d_off = w.read_off.to(torch.int64).contiguous()
d_out = torch.empty(2 * total + 64 * n + 4096, dtype=torch.uint8, device="cuda")
d_ooff = torch.zeros(n + 1, dtype=torch.int64, device="cuda")
d_st = torch.zeros(n, dtype=torch.uint8, device="cuda")
for _ in range(iters):
    ctr = t.correct_device(d_reads, d_off, total, d_out, d_ooff, d_st)
print("reads %d bases %d ms_correct %.1f ms_cov %.2f Mbp/s %.1f rounds %d launches %d" % (n, total, ctr["ms_correct"], ctr["ms_coverage"], total / 1e3 / ctr["ms_total"], ctr["rounds"], ctr["kernel_launches"]))
