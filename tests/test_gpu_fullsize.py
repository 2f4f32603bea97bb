"""GPU parity at the sizes BASELINE.json states: the CUDA path through the C ABI against the CPU
oracle on whole configurations, not scaled-down samples.

  config 1   1 000 transcripts, 10 000 reads, k=21 (the "bit-exact gate" config), every read, every counter
  config 5   k=30, 15 % error, 10 000 reads: gardening with ties, frontier > 50 aborts, cycles, 500 b borders
  k=31       the flagged extension beyond the reference's -k clamp: oracle with the clamp lifted
  config 2   the 38 M-entry table and ONE FULL 131 072-read batch (the bench's step) against the oracle
  config 3   the same table with the junction dump, 16 384 reads

The workloads of configs 2/3 are generated on the GPU (seconds); the oracle corrects them on the host cores."""
import hashlib
import json
import os

import numpy as np
import pytest

from conftest import Case, have_gpu

pytestmark = pytest.mark.gpu

SHARED = ["lookups_seg", "lookups_deg", "lookups_walk", "steps_inner", "steps_border", "frontier_sum", "cells_nw",
          "cells_lcs", "cells_ovl", "cells_xdrop", "gaps", "gaps_bridged", "gap_attempts", "borders",
          "borders_corrected", "ev_gardening", "ev_bridge", "ev_edge", "ev_cycle", "bases_out"]
PINS = os.path.join(os.path.dirname(__file__), "golden", "oracle_sha256.json")


@pytest.fixture(scope="module")
def api():
    if not have_gpu():
        pytest.fail("GPU tests selected but no CUDA device is visible")
    from talc_b200 import api as _api
    return _api


def _digest(out, off, status):
    """sha256 over what the three graded files are made of: the corrected bases, the read boundaries, the log."""
    h = hashlib.sha256()
    h.update(np.ascontiguousarray(out, dtype=np.uint8).tobytes())
    h.update(np.ascontiguousarray(off, dtype=np.uint64).tobytes())
    h.update(np.ascontiguousarray(status, dtype=np.uint8).tobytes())
    return h.hexdigest()


def _input_digest(case):
    h = hashlib.sha256()
    for a in (case.keys, case.counts, case.reads, case.off) + ((case.jkeys, case.jcounts) if case.jkeys is not None else ()):
        h.update(np.ascontiguousarray(a).tobytes())
    return h.hexdigest()


def _check(case, out, off, status, ctr, pin_name=None):
    n = len(case.off) - 1
    assert np.array_equal(status, case.o_status), "per-read status (failed-read log) differs"
    if not (np.array_equal(off, case.o_off) and np.array_equal(out, case.o_out)):
        bad = [r for r in range(n) if out[int(off[r]):int(off[r + 1])].tobytes() != case.oracle_read(r)]
        raise AssertionError("corrected sequence differs for %d reads, first %s" % (len(bad), bad[:10]))
    diff = {k: (case.o_ctr[k], ctr[k]) for k in SHARED if case.o_ctr[k] != ctr[k]}
    assert not diff, "algorithmic counters differ: %s" % diff
    assert ctr["reads_overflow"] == 0
    if pin_name:
        # the committed digest holds for the workload generated where the pin was made (torch CPU generator); if this
        # machine's libm rounds one transcript length differently the inputs differ and only the live oracle applies
        pin = json.load(open(PINS))[pin_name]
        if _input_digest(case) == pin["input_sha256"]:
            assert _digest(out, off, status) == pin["sha256"], "output differs from the committed oracle digest"


def test_config1_at_its_stated_size(api):
    case = Case(1, 1.0, 10000, threads=os.cpu_count() or 8)
    t = api.Talc(api.default_params(case.cfg.k))
    t.load_packed(case.keys, case.counts)
    out, off, st, ctr = t.correct(case.reads, case.off)
    _check(case, out, off, st, ctr, "config1")


def test_config5_at_its_stated_size(api):
    case = Case(5, 1.0, 10000, threads=os.cpu_count() or 8)
    assert case.o_ctr["ev_gardening"] > 1000 and case.o_ctr["ev_frontier_over50"] > 1000 and case.o_ctr["ev_cycle"] > 0
    t = api.Talc(api.default_params(case.cfg.k))
    t.load_packed(case.keys, case.counts)
    out, off, st, ctr = t.correct(case.reads, case.off)
    _check(case, out, off, st, ctr, "config5")


def test_k31_extension(api):
    """BASELINE config 5 names k=31; the reference's parser clamps -k to 30 (main.cpp:115).  The library
    accepts 31 (62 key bits) and must agree with the oracle run with the clamp lifted."""
    case = Case(5, 0.3, 1500, threads=os.cpu_count() or 8, k=31)
    assert case.o_ctr["gaps_bridged"] > 0 and case.o_ctr["ev_gardening"] > 0
    t = api.Talc(api.default_params(31))
    t.load_packed(case.keys, case.counts)
    out, off, st, ctr = t.correct(case.reads, case.off)
    _check(case, out, off, st, ctr, "k31")


_full = {}


def _full_workload():
    """Config 2 = config 3 minus the junction dump: one transcriptome, one count table, generated once on the GPU."""
    if not _full:
        import torch
        from talc_b200 import synth
        cfg = synth.baseline_config(2, 1.0)
        dev = "cuda:0"
        tr = synth.make_transcriptome(cfg, dev)
        keys, counts, jk, jc = synth.make_counts(cfg, tr, dev)
        reads, roff = synth.make_reads(cfg, tr, 131072, dev, seed_offset=3)
        del tr
        _full.update(cfg=cfg, keys=keys.cpu().numpy().astype(np.uint64), counts=counts.cpu().numpy().astype(np.int64),
                     jk=jk.cpu().numpy().astype(np.uint64), jc=jc.cpu().numpy().astype(np.int64), reads=reads.cpu().numpy(),
                     roff=roff.cpu().numpy().astype(np.uint64))
        del keys, counts, jk, jc, reads, roff
        torch.cuda.empty_cache()
    return _full


def _full_table_case(api, junctions, n_reads):
    from oracle import pyoracle as po
    if os.environ.get("TALC_SKIP_FULL_TABLE"):
        pytest.skip("TALC_SKIP_FULL_TABLE is set")
    w = _full_workload()
    cfg, keys, counts = w["cfg"], w["keys"], w["counts"]
    jk = w["jk"] if junctions else None
    jc = w["jc"] if junctions else None
    roff = w["roff"][: n_reads + 1]
    reads = w["reads"][: int(roff[-1])]
    t = api.Talc(api.default_params(cfg.k))
    t.load_packed(keys, counts, jk, jc)
    out, off, st, ctr = t.correct(reads, roff)
    ot = po.OracleTable(po.make_params(k=cfg.k)).build_packed(keys, counts, jk, jc)
    assert t.table_info()["entries"] == ot.size() > 25_000_000
    o_out, o_off, o_st, o_ctr, _ = ot.correct(reads, roff, threads=os.cpu_count() or 8)
    assert np.array_equal(st, o_st), "per-read status differs"
    if not (np.array_equal(off, o_off) and np.array_equal(out, o_out)):
        bad = [r for r in range(n_reads)
               if out[int(off[r]):int(off[r + 1])].tobytes() != o_out[int(o_off[r]):int(o_off[r + 1])].tobytes()]
        raise AssertionError("corrected sequence differs for %d reads, first %s" % (len(bad), bad[:10]))
    diff = {k: (o_ctr[k], ctr[k]) for k in SHARED if o_ctr[k] != ctr[k]}
    assert not diff, "algorithmic counters differ: %s" % diff
    return ctr


def test_config2_full_table_one_full_batch(api):
    """The bench's step: 131 072 reads in flight against the 38 M-entry table (L2-missing probes, cost-ordered
    scheduling, second-tier reads), every read compared with the oracle."""
    ctr = _full_table_case(api, False, 131072)
    assert ctr["reads_ok"] > 120000 and ctr["gaps_bridged"] > 1_000_000


def test_config3_full_table_with_junctions(api):
    ctr = _full_table_case(api, True, 16384)
    assert ctr["reads_ok"] > 15000
