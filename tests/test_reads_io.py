"""Host logic of the streamed `talc` command line without a GPU: the incremental FASTA / FASTQ reader and the FASTA
formatter of talc_b200/csrc/host/reads_io.hpp (what io.cpp:26-75 does in the reference) against
a plain Python restatement, over batch sizes that cut the input at every possible place."""
import ctypes as C
import os
import random
import subprocess

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "hostemu", "reads_io_test.cpp")
HDR = os.path.join(HERE, "..", "talc_b200", "csrc", "host", "reads_io.hpp")
LIB = os.path.join(HERE, "hostemu", "_build", "libreads_io_test.so")


@pytest.fixture(scope="module")
def rio():
    os.makedirs(os.path.dirname(LIB), exist_ok=True)
    if not os.path.exists(LIB) or any(os.path.getmtime(d) > os.path.getmtime(LIB) for d in (SRC, HDR)):
        subprocess.check_call(["/usr/bin/g++", "-O2", "-std=c++17", "-pthread", "-fPIC", "-shared", SRC, "-o", LIB])
    L = C.CDLL(LIB)
    L.rio_parse.restype = C.c_long
    L.rio_parse.argtypes = [C.c_char_p, C.c_long, C.c_long, C.c_char_p, C.c_long, C.c_void_p, C.c_long, C.c_void_p, C.c_long,
                            C.c_void_p, C.c_long, C.POINTER(C.c_long)]
    L.rio_format.restype = C.c_long
    L.rio_format.argtypes = [C.c_char_p, C.c_long, C.c_void_p, C.c_void_p, C.c_int, C.c_char_p, C.c_long]
    return L


def parse(L, path, batch_reads, batch_bases=1 << 40):
    size = (os.path.getsize(path) if os.path.exists(path) else 0) + 64
    ids = C.create_string_buffer(size)
    bases = np.zeros(size, dtype=np.uint8)
    offs = np.zeros(size, dtype=np.uint64)
    bs = np.zeros(size, dtype=np.int64)
    nb = C.c_long(0)
    n = L.rio_parse(path.encode(), batch_reads, batch_bases, ids, size, bases.ctypes.data, size, offs.ctypes.data, size,
                    bs.ctypes.data, size, C.byref(nb))
    if n < 0:
        return n, None, None, None
    names = ids.value.decode().split("\n")[:n] if n else []
    seqs = [bases[int(offs[i]):int(offs[i + 1])].tobytes().decode() for i in range(n)]
    return n, names, seqs, [int(x) for x in bs[:nb.value]]


def rand_seq(rng, n, alphabet="ACGT"):
    return "".join(rng.choice(alphabet) for _ in range(n))


def make_records(rng, n):
    recs = []
    for i in range(n):
        ln = rng.choice([0, 1, 5, 69, 70, 71, 140, 333]) if rng.random() < 0.5 else rng.randrange(0, 600)
        recs.append(("read_%d some description %d" % (i, rng.randrange(1000)), rand_seq(rng, ln, "ACGTNacgtn")))
    return recs


def write_fasta(path, recs, width, eol="\n", final_newline=True, blank_lines=False):
    with open(path, "w", newline="") as f:
        out = []
        for name, s in recs:
            out.append(">" + name)
            if width:
                out.extend(s[j:j + width] for j in range(0, len(s), width))
            else:
                out.append(s)
            if blank_lines:
                out.append("")
        txt = eol.join(out) + (eol if final_newline else "")
        f.write(txt)


def write_fastq(path, recs, eol="\n", final_newline=True):
    with open(path, "w", newline="") as f:
        out = []
        for name, s in recs:
            out += ["@" + name, s, "+", "I" * len(s)]
        f.write(eol.join(out) + (eol if final_newline else ""))


@pytest.mark.parametrize("eol", ["\n", "\r\n"])
def test_fasta_reader_any_batching(rio, tmp_path, eol):
    rng = random.Random(5)
    recs = make_records(rng, 57)
    for width, final_newline, blank in ((0, True, False), (60, True, False), (7, False, False), (80, True, True)):
        p = str(tmp_path / "r.fa")
        write_fasta(p, recs, width, eol, final_newline, blank)
        for batch_reads in (1, 2, 3, 7, 56, 57, 58, 1000):
            n, names, seqs, sizes = parse(rio, p, batch_reads)
            assert n == len(recs)
            assert names == [r[0] for r in recs]
            assert seqs == [r[1] for r in recs]
            assert sum(sizes) == n and all(0 < x <= batch_reads for x in sizes)
        # a batch also ends once it holds batch_bases bases (checked at record boundaries)
        n, names, seqs, sizes = parse(rio, p, 1000, 500)
        assert n == len(recs) and seqs == [r[1] for r in recs] and len(sizes) > 3


@pytest.mark.parametrize("eol", ["\n", "\r\n"])
def test_fastq_reader_any_batching(rio, tmp_path, eol):
    rng = random.Random(9)
    recs = [(nm, s) for nm, s in make_records(rng, 41)]
    p = str(tmp_path / "r.fq")
    for final_newline in (True, False):
        write_fastq(p, recs, eol, final_newline)
        for batch_reads in (1, 4, 40, 41, 500):
            n, names, seqs, sizes = parse(rio, p, batch_reads)
            assert n == len(recs) and names == [r[0] for r in recs] and seqs == [r[1] for r in recs]
            assert sum(sizes) == n


def test_reader_long_lines_and_large_file(rio, tmp_path):
    """an unwrapped 40 Mbp record (longer than the 16 MB read buffer) between short ones, and a line that straddles the
    buffer boundary"""
    rng = random.Random(2)
    big = rand_seq(rng, 1000) * 40000
    recs = [("a", "ACGT" * 10), ("big", big), ("c", "TTGCA" * 3)]
    p = str(tmp_path / "big.fa")
    write_fasta(p, recs, 0)
    n, names, seqs, sizes = parse(rio, p, 2)
    assert n == 3 and names == ["a", "big", "c"] and seqs[1] == big and seqs[0] == recs[0][1] and seqs[2] == recs[2][1]
    many = [("r%d" % i, rand_seq(rng, 997)) for i in range(40000)]  # ~40 MB: several buffer refills
    write_fasta(p, many, 61)
    n, names, seqs, sizes = parse(rio, p, 8191)
    assert n == len(many) and seqs == [r[1] for r in many] and names[-1] == "r39999"


def test_reader_rejects_what_the_reference_rejects(rio, tmp_path):
    p = str(tmp_path / "bad.fa")
    open(p, "w").write(">x\nACGTXACGT\n")          # a letter outside ACGTN
    assert parse(rio, p, 10)[0] == -2
    open(p, "w").write("ACGT\n>x\nACGT\n")          # neither '>' nor '@' first
    assert parse(rio, p, 10)[0] == -2
    open(p, "w").write("@x\nACGT\n+\n")            # truncated FASTQ record
    assert parse(rio, p, 10)[0] == -2
    open(p, "w").write("")                          # empty input: no reads, no error
    assert parse(rio, p, 10)[0] == 0
    open(p, "w").write(">only a header")            # a header without sequence is a read of length 0
    n, names, seqs, _ = parse(rio, p, 10)
    assert n == 1 and names == ["only a header"] and seqs == [""]
    open(p, "w").write(">x\nAC GT\tNN\n")           # blanks inside a sequence line are dropped
    assert parse(rio, p, 10)[2] == ["ACGTNN"]
    assert parse(rio, str(tmp_path / "missing.fa"), 10)[0] == -1


@pytest.mark.parametrize("threads", [1, 3, 16])
def test_fasta_formatter(rio, threads):
    rng = random.Random(threads)
    n = 5000
    recs = [("id_%d x" % i, rand_seq(rng, rng.choice([0, 1, 69, 70, 71, 139, 140, 141, rng.randrange(0, 400)]))) for i in range(n)]
    out = np.frombuffer("".join(s for _, s in recs).encode(), dtype=np.uint8).copy()
    offs = np.zeros(n + 1, dtype=np.uint64)
    offs[1:] = np.cumsum([len(s) for _, s in recs])
    ids = "\n".join(nm for nm, _ in recs).encode() + b"\n"
    cap = len(out) * 2 + 64 * n + 64
    dst = C.create_string_buffer(cap)
    got = rio.rio_format(ids, n, out.ctypes.data, offs.ctypes.data, threads, dst, cap)
    assert got > 0
    want = []
    for nm, s in recs:  # io.cpp:50-75: header, then 70 columns per line; an empty sequence still gets its (empty) line
        want.append(">" + nm + "\n")
        if not s:
            want.append("\n")
        want.extend(s[j:j + 70] + "\n" for j in range(0, len(s), 70))
    assert dst.raw[:got].decode() == "".join(want)
