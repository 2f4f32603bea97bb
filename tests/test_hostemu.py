"""The product's device code (talc_b200/csrc/*.cuh, the TALC_HD parts) compiled with g++ and diffed against the
oracle on the CPU.  This is a debugging aid for the build container, which has no GPU; it is not a fallback --
the shipped library only runs on CUDA.  The GPU runs of the same comparisons are in test_gpu_parity.py."""
import os

import numpy as np
import pytest

import pyemu
from oracle import pyoracle as po

SHARED = ["lookups_seg", "lookups_deg", "lookups_walk", "steps_inner", "steps_border", "frontier_sum", "cells_nw",
          "cells_lcs", "cells_ovl", "cells_xdrop", "gaps", "gaps_bridged", "gap_attempts", "borders",
          "borders_corrected", "ev_gardening", "ev_bridge", "ev_edge", "ev_cycle", "bases_out"]
GOLD = os.path.join(os.path.dirname(__file__), "golden", "tiny_case.npz")


def _emu_correct(case_params, keys, counts, jkeys, jcounts, reads, off):
    et = pyemu.EmuTable(pyemu.params_from(case_params), keys, counts, jkeys, jcounts)
    out, ooff, st, ctr = et.correct(reads, off, arena_bytes=48 * 1024, wide=False)
    ov = np.nonzero(st == 250)[0]
    if len(ov):  # second tier for the reads whose small slice overflowed
        parts = [reads[int(off[r]):int(off[r + 1])] for r in ov]
        so = np.zeros(len(ov) + 1, dtype=np.uint64)
        so[1:] = np.cumsum([len(x) for x in parts])
        o2, f2, s2, c2 = et.correct(np.concatenate(parts), so, arena_bytes=32 << 20, wide=True)
        assert not (s2 == 250).any()
        pieces, m = [], {int(r): i for i, r in enumerate(ov)}
        for r in range(len(off) - 1):
            if r in m:
                i = m[r]
                pieces.append(o2[int(f2[i]):int(f2[i + 1])])
                st[r] = s2[i]
            else:
                pieces.append(out[int(ooff[r]):int(ooff[r + 1])])
        ooff = np.zeros(len(off), dtype=np.uint64)
        ooff[1:] = np.cumsum([len(x) for x in pieces])
        out = np.concatenate(pieces)
        for k in ctr:
            ctr[k] += c2[k]
    return out, ooff, st, ctr


@pytest.mark.parametrize("name", ["c1", "c3", "c5"])
def test_device_code_on_host_matches_golden(name):
    g = np.load(GOLD)
    k = int(g[name + "_k"][0])
    usej = len(g[name + "_jkeys"]) > 0
    out, off, st, ctr = _emu_correct(po.make_params(k=k), g[name + "_keys"], g[name + "_counts"].astype(np.int64),
                                     g[name + "_jkeys"] if usej else None,
                                     g[name + "_jcounts"].astype(np.int64) if usej else None, g[name + "_reads"], g[name + "_off"])
    assert np.array_equal(st, g[name + "_status"])
    assert np.array_equal(off, g[name + "_ooff"]) and np.array_equal(out, g[name + "_out"])
    import json
    want = json.loads(bytes(g[name + "_ctr"]).decode())
    assert {k2: ctr[k2] for k2 in SHARED} == {k2: want[k2] for k2 in SHARED}


def test_device_code_on_host_matches_oracle_stress(case_c5):
    case = case_c5
    out, off, st, ctr = _emu_correct(case.op, case.keys, case.counts, None, None, case.reads, case.off)
    assert np.array_equal(st, case.o_status)
    assert np.array_equal(out, case.o_out) and np.array_equal(off, case.o_off)
    assert {k: ctr[k] for k in SHARED} == {k: case.o_ctr[k] for k in SHARED}
    assert case.o_ctr["ev_gardening"] > 0 and case.o_ctr["ev_sort_gt16"] > 0


def test_cycle_mode_switch_is_consistent(case_c5):
    """cycle_mode = 1 (exact first occurrence) must agree between the two implementations as well."""
    case = case_c5
    p1 = po.make_params(k=case.cfg.k, cycle_mode=1)
    ot = po.OracleTable(p1).build_packed(case.keys, case.counts)
    n = 60
    sub, so = case.reads[: int(case.off[n])], case.off[: n + 1]
    o_out, o_off, o_st, _, _ = ot.correct(sub, so, threads=4)
    out, off, st, _ = _emu_correct(p1, case.keys, case.counts, None, None, sub, so)
    assert np.array_equal(st, o_st) and np.array_equal(out, o_out)


def test_primitives_random():
    L = pyemu.lib()
    rng = np.random.default_rng(11)
    import ctypes as C
    for i in range(300):
        n, m = int(rng.integers(1, 300)), int(rng.integers(1, 300))
        a = bytes(rng.choice(list(b"ACGT"), n).tolist())
        b = bytearray(a[:m] if rng.random() < 0.7 else bytes(rng.choice(list(b"ACGT"), m).tolist()))
        for j in range(len(b)):
            if rng.random() < 0.1:
                b[j] = rng.choice(list(b"ACGT"))
        b = bytes(b) or b"A"
        if i % 9 == 0:
            a = a[: n // 2] + b"N" + a[n // 2:]
        for packed in (0, 1):
            assert L.emu_nw(a, b, packed) == po.nw(a, b)
            assert L.emu_lcs(a, b, packed) == po.lcs(a, b)
        assert L.emu_overlap(a, b) == po.overlap(a, b, True)
        for x in (-2, 0, 1, 5, 20):
            er, ec, ov = C.c_uint64(0), C.c_uint64(0), C.c_int(0)
            L.emu_xdrop(b, a, x, 1, C.byref(er), C.byref(ec), C.byref(ov))
            assert (er.value, ec.value) == po.xdrop(b, a, False, x) and ov.value == 0


def test_xdrop_register_band_mirror_matches_scalar():
    """csrc/xdrop.cuh keeps the X-drop band in registers, S diagonals per lane.  Its host mirror (same cell,
    window-update and end-position helpers, registers as flat arrays) must reproduce the cell-by-cell routine --
    end position AND number of cells visited -- for every S that admits the drop-off."""
    L = pyemu.lib()
    for seed, (n, max_len) in enumerate([(40000, 12), (40000, 60), (8000, 400)]):
        assert L.emu_xdrop_reg_fuzz(1000 + seed, n, max_len) == 0


def test_xdrop_landau_vishkin_characterisation_matches_scalar():
    """Groundwork for an incremental border scoring (what comes next): the X-drop's surviving cells are, per
    diagonal, the row interval that ends at the Landau-Vishkin furthest point of level X, with SeqAn's two boundary
    quirks; replaying the window rules on that set reproduces end position, end score and the number of cells
    visited of the cell-by-cell routine."""
    L = pyemu.lib()
    for mode in (0, 1, 2):  # replay every anti-diagonal / jump over quiet stretches (any boundary, outermost only)
        L.emu_xdrop_lv_mode(mode)
        for seed, (n, max_len) in enumerate([(30000, 10), (20000, 40), (4000, 300)]):
            assert L.emu_xdrop_lv_fuzz(2000 + seed, n, max_len) == 0
        assert L.emu_xdrop_lv_fuzz_realistic(5, 300, 600, 100, 24) == 0


def test_std_sort_replica_matches_libstdcxx():
    L = pyemu.lib()
    rng = np.random.default_rng(5)
    for n in list(range(1, 40)) + [64, 100, 150, 207, 500, 1000]:
        for hi in (2, 3, 10, 10 ** 6):
            keys = np.ascontiguousarray(rng.integers(0, hi, n), dtype=np.int64)
            perm = np.zeros(n, dtype=np.uint32)
            L.emu_std_sort_perm(keys.ctypes.data, n, perm.ctypes.data)
            assert np.array_equal(perm, po.std_sort_perm(keys)), (n, hi)
    for keys in (np.arange(300)[::-1].copy(), np.zeros(100, dtype=np.int64), np.arange(257)):
        keys = np.ascontiguousarray(keys, dtype=np.int64)
        perm = np.zeros(len(keys), dtype=np.uint32)
        L.emu_std_sort_perm(keys.ctypes.data, len(keys), perm.ctypes.data)
        assert np.array_equal(perm, po.std_sort_perm(keys))


def test_tagging_matches_oracle_exhaustively():
    L = pyemu.lib()
    import ctypes as C
    p = po.make_params(k=21)
    ep = pyemu.params_from(p)
    rng = np.random.default_rng(2)
    for _ in range(4000):
        count = int(rng.choice([2, 3, 4, 7, 30, 80, 81, 400, 5000, 120000]))
        c4 = rng.choice([0, 1, 2, 3, 5, 9, 15, 25, 60, 390, 4000], 4).astype(np.uint32)
        l4 = rng.choice([0, 0, 0, 7], 4).astype(np.uint32)
        cx = bool(rng.integers(0, 2))
        n1, t1, d1 = po.tag_next_nodes(p, c4, l4, count, cx)
        t2 = np.zeros(4, dtype=np.int32)
        d2 = np.zeros(4, dtype=np.float64)
        n2 = L.emu_tag_next_nodes(C.byref(ep), c4.ctypes.data, l4.ctypes.data, count, int(cx), t2.ctypes.data, d2.ctypes.data)
        assert n1 == n2
        if n1:
            assert list(t1) == list(t2)
            for i in range(4):
                if t1[i] != 1:
                    assert d1[i] == d2[i]


def test_suspend_and_resume_is_result_neutral(case_c5):
    """The per-read program as a resumable state machine (correct.cuh frames): a read that yields at every long walk
    and is resumed from its saved frames gives the bytes, status and counters of the run that never yields."""
    import pyemu
    from oracle import pyoracle as po
    g = np.load(GOLD)
    for name in ("c1", "c3", "c5"):
        k = int(g[name + "_k"][0])
        usej = len(g[name + "_jkeys"]) > 0
        et = pyemu.EmuTable(pyemu.params_from(po.make_params(k=k)), g[name + "_keys"], g[name + "_counts"].astype(np.int64),
                            g[name + "_jkeys"] if usej else None, g[name + "_jcounts"].astype(np.int64) if usej else None)
        y0 = pyemu.lib().emu_yields()
        a = et.correct(g[name + "_reads"], g[name + "_off"], arena_bytes=1 << 20, wide=True, split=True)
        assert pyemu.lib().emu_yields() > y0, "the split run must actually yield"
        b = et.correct(g[name + "_reads"], g[name + "_off"], arena_bytes=1 << 20, wide=True)
        d = et.correct(g[name + "_reads"], g[name + "_off"], arena_bytes=1 << 20, wide=True, state_machine=True)
        assert np.array_equal(d[0], b[0]) and np.array_equal(d[2], b[2]) and d[3] == b[3]  # state machines, no yield
        c = et.correct(g[name + "_reads"], g[name + "_off"], arena_bytes=1 << 20, wide=True, split=True, pause_every=1)
        assert np.array_equal(c[0], b[0]) and np.array_equal(c[2], b[2]) and c[3] == b[3]  # paused after every general step
        assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1]) and np.array_equal(a[2], b[2]) and a[3] == b[3]
        assert np.array_equal(a[0], g[name + "_out"]) and np.array_equal(a[2], g[name + "_status"])
    case = case_c5
    et = pyemu.EmuTable(pyemu.params_from(case.op), case.keys, case.counts)
    out, off, st, ctr = et.correct(case.reads, case.off, arena_bytes=4 << 20, wide=True, split=True, pause_every=3)
    assert np.array_equal(st, case.o_status) and np.array_equal(off, case.o_off) and np.array_equal(out, case.o_out)
    for k2 in ("lookups_walk", "steps_inner", "steps_border", "cells_xdrop", "ev_gardening", "ev_cycle"):
        assert ctr[k2] == case.o_ctr[k2], k2


def test_two_contexts_interleaved_on_one_thread():
    """Two read contexts multiplexed on one thread (start A, start B, resume whichever has been walked, contexts reused
    from read to read): no state of a suspended read lives outside its Corrector and its arena."""
    g = np.load(GOLD)
    for name in ("c1", "c3", "c5"):
        k = int(g[name + "_k"][0])
        usej = len(g[name + "_jkeys"]) > 0
        et = pyemu.EmuTable(pyemu.params_from(po.make_params(k=k)), g[name + "_keys"], g[name + "_counts"].astype(np.int64),
                            g[name + "_jkeys"] if usej else None, g[name + "_jcounts"].astype(np.int64) if usej else None)
        out, off, st, ctr = pyemu.correct_interleaved(et, g[name + "_reads"], g[name + "_off"])
        assert np.array_equal(out, g[name + "_out"]) and np.array_equal(st, g[name + "_status"])
