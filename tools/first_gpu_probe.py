"""Quick GPU probe used during bring-up: times the hot path on a mid-size synthetic workload."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from talc_b200 import api, synth

scale = float(sys.argv[1]) if len(sys.argv) > 1 else 0.05
cfg = synth.baseline_config(2, scale)
t0 = time.time()
w = synth.make_workload(cfg, device="cuda")
torch.cuda.synchronize()
print("generated: keys %d reads %d bases %d in %.1fs" % (w.keys.numel(), w.n_reads(), w.total_bases(), time.time() - t0), flush=True)
t = api.Talc(api.default_params(cfg.k))
t0 = time.time()
nk = t.load_packed(w.keys.cpu().numpy().astype(np.uint64), w.counts.cpu().numpy())
print("table: kept %d in %.1fs" % (nk, time.time() - t0), t.table_info(), flush=True)
d_reads = w.reads.contiguous()
d_off = w.read_off.to(torch.int64).contiguous()
n = w.n_reads()
total = w.total_bases()
d_out = torch.empty(2 * total + 64 * n + 4096, dtype=torch.uint8, device="cuda")
d_ooff = torch.zeros(n + 1, dtype=torch.int64, device="cuda")
d_st = torch.zeros(n, dtype=torch.uint8, device="cuda")
for tier1 in (1 << 20,):
    for it in range(3):
        t0 = time.time()
        ctr = t.correct_device(d_reads, d_off, total, d_out, d_ooff, d_st)
        dt = time.time() - t0
        print("tier1=%dK it%d wall %.3fs  Mbp/s %.1f  ms: cov %.1f correct %.1f tier2 %.1f gather %.1f total %.1f  second-tier reads %d  steps %d lookups_walk %d" % (
            tier1 // 1024, it, dt, total / 1e6 / dt, ctr["ms_coverage"], ctr["ms_correct"], ctr["ms_correct_tier2"], ctr["ms_gather"], ctr["ms_total"],
            ctr["reads_second_tier"], ctr["steps_inner"] + ctr["steps_border"], ctr["lookups_walk"]), flush=True)
st = d_st.cpu().numpy()
print("status histogram", np.bincount(st, minlength=4))
print({k: ctr[k] for k in ("gaps", "gaps_bridged", "borders", "borders_corrected", "cells_nw", "cells_lcs", "cells_xdrop", "bases_out")})
