"""A/B of library builds on one box: times talc_correct_batch_device of a given libtalc_b200.so (any commit: only the
entry points every version has are bound) on the bench workload.  usage: python tools/time_lib.py lib1.so lib2.so ..."""
import ctypes as C, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from talc_b200 import synth

class P(C.Structure):
    _fields_ = [("K", C.c_uint32), ("min_count", C.c_uint32), ("window_size", C.c_uint32), ("max_nb_branches", C.c_uint32),
                ("alpha", C.c_double), ("sr_error_rate", C.c_double), ("min_inner_score", C.c_double),
                ("min_border_score", C.c_double), ("cycle_mode", C.c_int32), ("q11_zero_init", C.c_int32)]

scale = float(os.environ.get("TALC_AB_SCALE", "1.0"))
B = int(os.environ.get("TALC_AB_READS", "131072"))
cfg = synth.baseline_config(int(os.environ.get("TALC_AB_CONFIG", "2")), scale)
dev = "cuda:0"
tr = synth.make_transcriptome(cfg, dev)
keys, counts, jk, jc = synth.make_counts(cfg, tr, dev)
reads, roff = synth.make_reads(cfg, tr, B * 2, dev, seed_offset=3)
roff = roff.to(torch.int64)
hk, hc = keys.cpu().numpy().astype(np.uint64), counts.cpu().numpy().astype(np.int64)
del tr, keys, counts
vp = C.c_void_p
for path in sys.argv[1:]:
    L = C.CDLL(os.path.abspath(path))
    L.talc_params_default.argtypes = [C.POINTER(P), C.c_uint32]
    L.talc_ctx_create.argtypes = [C.POINTER(P), C.c_int, C.POINTER(vp)]
    L.talc_table_load_packed.argtypes = [vp, vp, vp, C.c_uint64, vp, vp, C.c_uint64, C.c_int, C.POINTER(C.c_uint64)]
    L.talc_correct_batch_device.argtypes = [vp, vp, vp, C.c_uint32, C.c_uint64, vp, C.c_uint64, vp, vp, vp]
    L.talc_ctx_destroy.argtypes = [vp]
    p = P(); L.talc_params_default(C.byref(p), cfg.k)
    h = vp(); assert L.talc_ctx_create(C.byref(p), 0, C.byref(h)) == 0
    nk = C.c_uint64(0)
    assert L.talc_table_load_packed(h, hk.ctypes.data, hc.ctypes.data, len(hk), None, None, 0, 0, C.byref(nk)) == 0
    times = []
    for it in range(4):
        i = it % 2
        lo, hi = i * B, (i + 1) * B
        b0, b1 = int(roff[lo]), int(roff[hi])
        r = reads[b0:b1].contiguous(); o = (roff[lo:hi + 1] - roff[lo]).contiguous()
        out = torch.empty(2 * (b1 - b0) + 64 * B + 4096, dtype=torch.uint8, device=dev)
        oo = torch.zeros(B + 1, dtype=torch.int64, device=dev); st = torch.zeros(B, dtype=torch.uint8, device=dev)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        rc = L.talc_correct_batch_device(h, r.data_ptr(), o.data_ptr(), B, b1 - b0, out.data_ptr(), out.numel(), oo.data_ptr(), st.data_ptr(), None)
        dt = time.perf_counter() - t0
        assert rc == 0
        times.append(dt * 1e3)
    print("%-40s ms per batch: %s  (Mbp/s %.1f)" % (os.path.basename(path), " ".join("%.1f" % t for t in times), (b1 - b0) / 1e3 / min(times[1:])), flush=True)
    L.talc_ctx_destroy(h)
