"""CPU tests of the oracle: hand-worked known answers for every primitive it restates, and the committed
golden vectors (tests/golden/tiny_case.npz, made by tools/make_golden.py).  The reference itself ships no
tests or vectors and cannot be built here (SeqAn2 missing) -- see oracle/talc_oracle.hpp 'PARITY UNPINNED'."""
import json
import os

import numpy as np
import pytest

from oracle import pyoracle as po

GOLD = os.path.join(os.path.dirname(__file__), "golden", "tiny_case.npz")


def test_nw_lcs_known_answers():
    assert po.nw(b"ACGT", b"ACGT") == 0
    assert po.nw(b"ACGT", b"AGT") == -1          # one deletion
    assert po.nw(b"ACGT", b"TGCA") == -4
    assert po.nw(b"AAAA", b"A") == -3
    assert po.nw(b"ACGTN", b"ACGTN") == 0          # N matches itself
    assert po.nw(b"ACGTN", b"ACGTA") == -1
    assert po.lcs(b"ACGT", b"AGT") == 3
    assert po.lcs(b"AAAA", b"TTTT") == 0
    assert po.lcs(b"ACGTACGT", b"CGTTACG") == 6


def test_overlap_score_known_answers():
    # RIGHT = leading gaps free in both sequences, score = bottom-right cell (Trail.cpp:170-172)
    assert po.overlap(b"ACGT", b"ACGT", True) == 16
    assert po.overlap(b"TTACGT", b"ACGT", True) == 16      # free leading gap in the candidate
    assert po.overlap(b"ACGT", b"ACGA", True) == 9         # 3 matches, 1 mismatch: 12 - 3
    assert po.overlap(b"ACGTTT", b"ACGT", True) == 12      # trailing gap of 2 is charged: 16 - 4
    # LEFT = trailing gaps free, leading gaps charged, score = max over last row / column (Trail.cpp:164-167)
    assert po.overlap(b"ACGTTT", b"ACGT", False) == 16
    assert po.overlap(b"TTACGT", b"ACGT", False) == 12
    # reversal symmetry used by the device code: LEFT(ref, cand) == RIGHT(rev ref, rev cand)
    rng = np.random.default_rng(0)
    for _ in range(50):
        a = bytes(rng.choice(list(b"ACGT"), int(rng.integers(1, 60))).tolist())
        b = bytes(rng.choice(list(b"ACGT"), int(rng.integers(1, 60))).tolist())
        assert po.overlap(a, b, False) == po.overlap(a[::-1], b[::-1], True)


def test_xdrop_known_answers():
    # identical segments extend to both ends whatever X >= 0 is
    assert po.xdrop(b"ACGTACGT", b"ACGTACGT", False, 0) == (8, 8)
    # X = 0: the extension stops at the first mismatch
    assert po.xdrop(b"ACGTTCGT", b"ACGTACGT", False, 0) == (4, 4)
    # X = 1 lets one mismatch through, the identical tail then reaches both ends
    assert po.xdrop(b"ACGTTCGT", b"ACGTACGT", False, 1) == (8, 8)
    # negative X: nothing extends (Q20 passes negative values)
    assert po.xdrop(b"ACGT", b"ACGT", False, -1) == (0, 0)
    # an empty side returns immediately
    assert po.xdrop(b"", b"ACGT", False, 5) == (0, 0)
    # EXTEND_LEFT consumes the prefixes from their ends: same as EXTEND_RIGHT on reversed strings
    rng = np.random.default_rng(1)
    for _ in range(100):
        a = bytes(rng.choice(list(b"ACGT"), int(rng.integers(1, 80))).tolist())
        b = bytearray(a)
        for i in range(len(b)):
            if rng.random() < 0.1:
                b[i] = rng.choice(list(b"ACGT"))
        b = bytes(b)[: int(rng.integers(1, len(a) + 1))]
        for x in (0, 1, 3, 7):
            assert po.xdrop(b, a, True, x) == po.xdrop(b[::-1], a[::-1], False, x)


def test_seed_extension_semantics():
    k = 5
    ref = b"ACGTAGGCTTAACC"
    # a perfect copy extends over everything; seed (0,0,K-1,K-1) (Q18) -> extension lengths include K-1
    re_, ce_, pos, score, stop = po.seed_extend(ref, ref, 2, True, k)
    assert (re_, ce_, pos, score, stop) == (len(ref), len(ref), len(ref), 0.0, False)
    # no extension possible with X < 0: both extensions stay at K-1 < K -> score = -xdrop, stop (Q19)
    re_, ce_, pos, score, stop = po.seed_extend(ref, ref, -3, True, k)
    assert (re_, ce_, score, stop) == (k - 1, k - 1, 3.0, True)
    # LEFT: the seed is the last K bases, so an unextended pair has length K and is scored (no stop)
    re_, ce_, pos, score, stop = po.seed_extend(ref, ref, -3, False, k)
    assert (re_, ce_, pos, score, stop) == (k, k, len(ref) - k, 0.0, False)


def test_horspool_modes():
    hay = b"ACGTACGTACGTAAAC"
    # exact first occurrence
    assert po.horspool(hay, b"CGTA", 1) == 1
    assert po.horspool(hay, b"TTTT", 1) == -1
    # SeqAn-2 mixed-alphabet Horspool: the shift is always |needle| -> only windows 0, K, 2K, ... are seen
    assert po.horspool(hay, b"CGTA", 0) == -1
    assert po.horspool(hay, b"ACGT", 0) == 0
    assert po.horspool(hay, b"AAAC", 0) == 12


def test_count_model_known_values():
    a = 2.57
    # cc <= 3, upper: next <= (cc+.5) + a*sqrt(cc+.5);  cc = 2 -> 2.5 + 2.57*1.5811 = 6.563
    assert po.lib().orc_is_expected_model(6, 2, a, 0) == 1
    assert po.lib().orc_is_expected_model(7, 2, a, 0) == 0
    # cc > 3, upper: (a/2 + sqrt(cc+.96))^2 ; cc = 30 -> (1.285 + 5.564)^2 = 46.9
    assert po.lib().orc_is_expected_model(46, 30, a, 0) == 1
    assert po.lib().orc_is_expected_model(47, 30, a, 0) == 0
    # cc > 3, lower: (a/2 - sqrt(cc+.02))^2 ; cc = 30 -> (1.285 - 5.479)^2 = 17.59
    assert po.lib().orc_is_expected_model(18, 30, a, 1) == 1
    assert po.lib().orc_is_expected_model(17, 30, a, 1) == 0
    # cc = 0, lower bound is sqrt(-0.5) = NaN -> never expected
    assert po.lib().orc_is_expected_model(5, 0, a, 1) == 0


def test_tag_next_nodes_rules():
    p = po.make_params(k=21)
    EXP, UNEXP, BRK = 0, 1, 7
    # dead end: no successor reaches MIN_COUNT -> no tags at all
    n, tags, _ = po.tag_next_nodes(p, [0, 1, 0, 1], [0, 0, 0, 0], 30, False)
    assert n == 0
    # a single admissible successor is EXPECTED whatever its count (counter == 1)
    n, tags, dist = po.tag_next_nodes(p, [0, 2, 0, 0], [0, 0, 0, 0], 400, False)
    assert n == 4 and list(tags) == [UNEXP, EXP, UNEXP, UNEXP]
    assert dist[1] == pytest.approx(abs(400 - 2) / 20.0)
    # lambda = (int)(count*0.025) < MIN_COUNT: every unexpected admissible successor is a BREAKPOINT;
    # a lone BREAKPOINT without any EXPECTED is promoted to EXPECTED
    n, tags, _ = po.tag_next_nodes(p, [0, 2, 0, 0], [0, 0, 0, 0], 30, False)
    assert list(tags) == [UNEXP, EXP, UNEXP, UNEXP]
    # count = 400: lambda = 10, noise upper bound (1.285+sqrt(10.96))^2 = 21.1
    n, tags, _ = po.tag_next_nodes(p, [390, 15, 0, 0], [0, 0, 0, 0], 400, False)
    assert list(tags) == [EXP, UNEXP, UNEXP, UNEXP]          # 15 is within noise -> UNEXPECTED
    n, tags, _ = po.tag_next_nodes(p, [390, 25, 0, 0], [0, 0, 0, 0], 400, False)
    assert list(tags) == [EXP, BRK, UNEXP, UNEXP]            # 25 is above noise -> BREAKPOINT
    n, tags, _ = po.tag_next_nodes(p, [390, 15, 0, 0], [0, 3, 0, 0], 400, False)
    assert list(tags) == [EXP, BRK, UNEXP, UNEXP]            # junction colour rescues it
    # one EXPECTED + several UNEXPECTED whose SUM is above noise: the largest becomes a BREAKPOINT ...
    n, tags, _ = po.tag_next_nodes(p, [390, 12, 11, 0], [0, 0, 0, 0], 400, False)
    assert list(tags) == [EXP, BRK, UNEXP, UNEXP]
    # ... unless the frontier is already complex
    n, tags, _ = po.tag_next_nodes(p, [390, 12, 11, 0], [0, 0, 0, 0], 400, True)
    assert list(tags) == [EXP, UNEXP, UNEXP, UNEXP]


def test_table_build_semantics(tmp_path):
    p = po.make_params(k=5)
    dump = tmp_path / "d.dump"
    dump.write_text("ACGTA 5\nACGTA 9\nCCCCC 7\nGGGGA 1\nTTTTT 3\nbad_line\nAAAAA 4\nTACGT 6\nacgtc 8\n")
    junc = tmp_path / "j.dump"
    junc.write_text("ACGTA 2\nTACGT 4\nACGTA 3\nAAAAA 5\nCCCCC 10000\nACGTC 0\n")
    t = po.OracleTable(p).load_dump(str(dump), str(junc))
    assert t.size() == 6                         # GGGGA (count 1 < MIN) and the bad line are dropped
    # ACGTA and TACGT are reverse complements of each other.  Colour writes, in file order:
    #   line 1 (ACGTA 2): ACGTA=2, TACGT=2;  line 2 (TACGT 4): TACGT=4, ACGTA=4;  line 3 (ACGTA 3): ACGTA=3, TACGT=3
    assert t.lookup("ACGTA") == (True, 5, 3)     # count: first dump line wins (5, not 9); colour: last writer wins
    assert t.lookup("TACGT") == (True, 6, 3)
    assert t.lookup("CCCCC") == (True, 7, 0)     # 10000 is not < 10000; homopolymer decoloured anyway
    assert t.lookup("AAAAA") == (True, 4, 0)     # coloured by the junction file, then decoloured (utils.cpp:658-669)
    assert t.lookup("TTTTT") == (True, 3, 0)     # rc of AAAAA got 5, then decoloured
    assert t.lookup("ACGTC") == (True, 8, 0)     # lower case converts; colour 0 written
    assert t.lookup("GGGGA") == (False, 0, 0)


@pytest.mark.parametrize("name", ["c1", "c3", "c5"])
def test_golden_vectors(name):
    g = np.load(GOLD)
    k = int(g[name + "_k"][0])
    usej = len(g[name + "_jkeys"]) > 0
    t = po.OracleTable(po.make_params(k=k)).build_packed(g[name + "_keys"], g[name + "_counts"].astype(np.int64),
                                                         g[name + "_jkeys"] if usej else None,
                                                         g[name + "_jcounts"].astype(np.int64) if usej else None)
    out, off, st, ctr, _ = t.correct(g[name + "_reads"], g[name + "_off"], threads=2)
    assert np.array_equal(st, g[name + "_status"])
    assert np.array_equal(off, g[name + "_ooff"]) and np.array_equal(out, g[name + "_out"])
    want = json.loads(bytes(g[name + "_ctr"]).decode())
    assert {k2: ctr[k2] for k2 in want if k2 != "frontier_max"} == {k2: want[k2] for k2 in want if k2 != "frontier_max"}
    # thread count must not change anything
    out1, off1, st1, _, _ = t.correct(g[name + "_reads"], g[name + "_off"], threads=1)
    assert np.array_equal(out1, out) and np.array_equal(st1, st)


def test_cli_outputs(tmp_path):
    """The oracle's `talc` front end writes the four files of the reference."""
    import subprocess
    from talc_b200 import synth
    g = np.load(GOLD)
    k = int(g["c1_k"][0])
    import torch
    synth.write_dump(str(tmp_path / "sr.dump"), torch.from_numpy(g["c1_keys"].astype(np.int64)),
                     torch.from_numpy(g["c1_counts"].astype(np.int64)), k)
    synth.write_fasta(str(tmp_path / "reads.fa"), torch.from_numpy(g["c1_reads"]), torch.from_numpy(g["c1_off"].astype(np.int64)))
    pre = str(tmp_path / "o")
    rc = subprocess.call([po.BIN, str(tmp_path / "reads.fa"), "--SRCounts", str(tmp_path / "sr.dump"), "-k", str(k), "-o", pre,
                          "-t", "2", "--oracle-table", "hash"], stdout=subprocess.DEVNULL)
    assert rc == 0
    cfg = open(pre + ".config.txt").read().splitlines()
    assert cfg[0] == "TALC: Parameters used for sample: " + pre and cfg[6] == "KmerSize=%d" % k and cfg[10] == "MIN_INNER_SCORE=0.7"
    assert len(cfg) == 17 and cfg[13] == "ALPHA=2.57"
    assert open(pre + ".stats_basics.txt").read().startswith("read_name\traw_length\t")
    fa = open(pre + ".fa").read().split(">")[1:]
    assert len(fa) == len(g["c1_off"]) - 1
    for r, rec in enumerate(fa):
        lines = rec.split("\n")
        assert lines[0] == "read_%d" % r and all(len(x) <= 70 for x in lines[1:])
        want = bytes(g["c1_out"][int(g["c1_ooff"][r]):int(g["c1_ooff"][r + 1])]).decode()
        assert "".join(lines[1:]) == want
    assert subprocess.call([po.BIN, "x.fa", "-k", "31", "--SRCounts", "d"], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL) == 1
