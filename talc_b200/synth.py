"""Seeded synthetic workloads for TALC's correction path.

A simulated transcriptome (genes + isoforms that skip or swap an internal segment),
a Jellyfish-style k-mer count table drawn from a short-read sampling model, an optional
splice-junction k-mer list, and ONT-like long reads (substitution:insertion:deletion
= 4:3:3).  Everything is vectorised torch so that the 1k-transcript gate case runs on the
CPU in a second and the 200k-transcript / 1M-read bench case is generated on the GPU.

Short-read model (per k-mer occurrence, not per read): read starts are Poisson per
transcript position, a k-mer's count is the number of reads that cover it entirely
(sliding-window sum, so neighbouring counts are correlated like real coverage), thinned
by the probability that a read is error-free over the k-mer; erroneous k-mer variants are
spawned with small counts so that tips/bubbles below and just above MIN_COUNT exist.

k-mers are 2-bit packed into int64, first base in the most significant position
(A=0, C=1, G=2, T=3) -- the same packing the C ABI uses (include/talc_b200.h).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Optional

import numpy as np
import torch

BASES = np.frombuffer(b"ACGT", dtype=np.uint8)


@dataclass
class SynthConfig:
    seed: int = 1
    k: int = 21
    n_genes: int = 700
    n_transcripts: int = 1000
    n_reads: int = 10000
    read_error: float = 0.10          # long-read error rate, split 4:3:3
    sr_len: int = 100                 # short-read length
    sr_error: float = 0.005           # short-read substitution rate
    cov_mu: float = math.log(30.0)    # lognormal coverage
    cov_sigma: float = 1.0
    variant_rate: float = 0.4         # erroneous k-mer variants spawned per k-mer occurrence
    lowcomp_frac: float = 0.0         # fraction of isoform inserts that are (AC)n / poly-A
    iso_min: int = 50
    iso_max: int = 300
    high_count_kmers: int = 0         # a few k-mers forced to counts >= 100000 (p_MAX_IN_COUNT)
    len_mu: float = math.log(1500.0)
    len_sigma: float = 0.5
    len_min: int = 300
    len_max: int = 6000


# the five BASELINE.json configurations; sizes can be scaled down for tests
def baseline_config(i: int, scale: float = 1.0) -> SynthConfig:
    if i == 1:
        c = SynthConfig(seed=1, k=21, n_genes=700, n_transcripts=1000, n_reads=10000, high_count_kmers=3)
    elif i in (2, 3):
        # 8k genes x ~25 isoforms: ~12M gene k-mers + ~16M novel/junction k-mers + ~4M erroneous k-mers with
        # count >= 2  ->  ~30M kept entries, as BASELINE.json names for the transcriptome-scale configuration
        c = SynthConfig(seed=2, k=21, n_genes=8000, n_transcripts=200000, n_reads=1000000, high_count_kmers=5,
                        variant_rate=0.05)
    elif i == 4:
        c = SynthConfig(seed=2, k=21, n_genes=8000, n_transcripts=200000, n_reads=5000000, high_count_kmers=5,
                        variant_rate=0.05)
    elif i == 5:
        c = SynthConfig(seed=5, k=30, n_genes=250, n_transcripts=1000, n_reads=10000, read_error=0.15,
                        lowcomp_frac=0.5, high_count_kmers=3)
    else:
        raise ValueError(i)
    if scale != 1.0:
        c.n_genes = max(4, int(c.n_genes * scale))
        c.n_transcripts = max(c.n_genes + 2, int(c.n_transcripts * scale))
        c.n_reads = max(8, int(c.n_reads * scale))
    return c


@dataclass
class Workload:
    cfg: SynthConfig
    # transcriptome
    t_bases: torch.Tensor = None      # uint8 codes 0..3, all transcripts concatenated
    t_off: torch.Tensor = None        # int64 [T+1]
    t_cov: torch.Tensor = None        # float64 [T]
    # count table, in dump order (seeded shuffle), one line per distinct k-mer, counts >= 1
    keys: torch.Tensor = None         # int64 packed k-mers
    counts: torch.Tensor = None       # int64
    # junction list (k-mers straddling isoform breakpoints), counts capped at 9999
    jkeys: torch.Tensor = None
    jcounts: torch.Tensor = None
    # long reads: ASCII bytes concatenated + offsets
    reads: torch.Tensor = None        # uint8 ASCII
    read_off: torch.Tensor = None     # int64 [R+1]
    extra: dict = field(default_factory=dict)

    def n_reads(self) -> int:
        return int(self.read_off.numel() - 1)

    def total_bases(self) -> int:
        return int(self.read_off[-1])


def _gen(seed: int, device) -> torch.Generator:
    g = torch.Generator(device=device)
    g.manual_seed(int(seed))
    return g


def pack_kmers(bases: torch.Tensor, k: int) -> torch.Tensor:
    """All k-mers of a code array (0..3) as int64, one per start position (len-k+1)."""
    n = bases.numel()
    b = bases.to(torch.int64)
    out = torch.zeros(n - k + 1, dtype=torch.int64, device=bases.device)
    for j in range(k):
        out = (out << 2) | b[j:n - k + 1 + j]
    return out


def unpack_kmer(key: int, k: int) -> str:
    return "".join("ACGT"[(key >> (2 * (k - 1 - i))) & 3] for i in range(k))


def make_transcriptome(cfg: SynthConfig, device="cpu"):
    g = _gen(cfg.seed * 1000 + 1, device)
    G, T, K = cfg.n_genes, cfg.n_transcripts, cfg.k
    f64 = dict(dtype=torch.float64, device=device)
    glen = torch.exp(torch.randn(G, generator=g, **f64) * cfg.len_sigma + cfg.len_mu)
    glen = glen.clamp(cfg.len_min, cfg.len_max).to(torch.int64)
    goff = torch.zeros(G + 1, dtype=torch.int64, device=device)
    goff[1:] = torch.cumsum(glen, 0)
    gbases = torch.randint(0, 4, (int(goff[-1]),), generator=g, device=device, dtype=torch.uint8)

    parent = torch.cat([torch.arange(G, device=device),
                        torch.randint(0, G, (T - G,), generator=g, device=device)])
    plen = glen[parent]
    u = torch.rand(T, generator=g, **f64)
    kind = torch.zeros(T, dtype=torch.int64, device=device)          # 0 copy
    iso = torch.arange(T, device=device) >= G
    r = torch.rand(T, generator=g, **f64)
    kind[iso & (r < 0.5)] = 1                                        # 1 segment removed
    kind[iso & (r >= 0.5)] = 2                                       # 2 segment swapped for novel sequence
    lc = torch.rand(T, generator=g, **f64) < cfg.lowcomp_frac
    kind[iso & lc] = 3                                               # 3 segment swapped for a low-complexity repeat
    d = torch.randint(cfg.iso_min, cfg.iso_max + 1, (T,), generator=g, device=device)
    d = torch.minimum(d, plen // 3)
    d[kind == 0] = 0
    room = (plen - d - 2 * K - 2).clamp(min=1)
    a = K + 1 + (u * room.to(torch.float64)).to(torch.int64)        # segment start inside the parent
    ins = torch.zeros(T, dtype=torch.int64, device=device)
    ins[kind == 2] = d[kind == 2]
    lclen = torch.randint(25, 41, (T,), generator=g, device=device)
    ins[kind == 3] = lclen[kind == 3]
    lcstyle = torch.randint(0, 2, (T,), generator=g, device=device)  # 0 (AC)n, 1 poly-A
    tlen = plen - d + ins
    toff = torch.zeros(T + 1, dtype=torch.int64, device=device)
    toff[1:] = torch.cumsum(tlen, 0)
    N = int(toff[-1])
    tid = torch.repeat_interleave(torch.arange(T, device=device), tlen)
    pos = torch.arange(N, device=device) - toff[tid]
    a_t, ins_t, d_t = a[tid], ins[tid], d[tid]
    in_pre = pos < a_t
    in_ins = (~in_pre) & (pos < a_t + ins_t)
    src = goff[parent[tid]] + torch.where(in_pre, pos, pos - ins_t + d_t)
    src = src.clamp(max=gbases.numel() - 1)
    tb = gbases[src]
    novel = torch.randint(0, 4, (N,), generator=g, device=device, dtype=torch.uint8)
    k_t = kind[tid]
    rep = torch.where(lcstyle[tid] == 0, ((pos - a_t) & 1), torch.zeros_like(pos)).to(torch.uint8)  # A/C alternate, or A
    tb = torch.where(in_ins & (k_t == 2), novel, tb)
    tb = torch.where(in_ins & (k_t == 3), rep, tb)
    cov = torch.exp(torch.randn(T, generator=g, **f64) * cfg.cov_sigma + cfg.cov_mu)
    # junction boundaries per transcript (transcript coordinates); -1 when absent
    b1 = torch.where(kind > 0, a, torch.full_like(a, -1))
    b2 = torch.where(ins > 0, a + ins, torch.full_like(a, -1))
    return dict(t_bases=tb, t_off=toff, t_cov=cov, tid=tid, pos=pos, tlen=tlen, b1=b1, b2=b2)


def make_counts(cfg: SynthConfig, tr: dict, device="cpu"):
    """k-mer counts under the short-read sampling model; returns dump-ordered (keys, counts) and junctions."""
    g = _gen(cfg.seed * 1000 + 2, device)
    K, Lr = cfg.k, cfg.sr_len
    tb, toff, cov, tid, pos, tlen = tr["t_bases"], tr["t_off"], tr["t_cov"], tr["tid"], tr["pos"], tr["tlen"]
    N = tb.numel()
    T = tlen.numel()
    # read starts: Poisson(cov/Lr) at every position where a full read fits (short transcripts: anywhere)
    rate = (cov / Lr)[tid]
    fits = pos <= (tlen[tid] - Lr).clamp(min=0)
    starts = torch.poisson(torch.where(fits, rate, torch.zeros_like(rate)), generator=g)
    cs = torch.cumsum(starts, 0)
    cs = torch.cat([torch.zeros(1, dtype=cs.dtype, device=device), cs])
    # k-mer at transcript position i is covered by reads starting in [i-(Lr-K), i]
    lo = (pos - (Lr - K)).clamp(min=0) + toff[tid]
    hi = pos + toff[tid]
    depth = cs[hi + 1] - cs[lo]
    valid = pos <= tlen[tid] - K
    p_ok = (1.0 - cfg.sr_error) ** K
    good = torch.binomial(depth, torch.full_like(depth, p_ok), generator=g)
    # packed k-mers over the concatenated transcripts (invalid windows masked out below)
    allk = torch.zeros(N, dtype=torch.int64, device=device)
    kk = pack_kmers(tb, K)
    allk[: kk.numel()] = kk
    del kk
    vk = allk[valid]
    vc = good[valid].to(torch.int64)
    # erroneous variants: one substitution at a random offset, small counts
    nv = int(cfg.variant_rate * vk.numel())
    if nv > 0:
        pick = torch.randint(0, vk.numel(), (nv,), generator=g, device=device)
        off = torch.randint(0, K, (nv,), generator=g, device=device)
        delta = torch.randint(1, 4, (nv,), generator=g, device=device)
        shift = 2 * (K - 1 - off)
        base = (vk[pick] >> shift) & 3
        newb = (base + delta) & 3
        var = (vk[pick] & ~(torch.full_like(shift, 3) << shift)) | (newb << shift)
        varc = 1 + torch.poisson(torch.full((nv,), 0.35, dtype=torch.float64, device=device), generator=g).to(torch.int64)
        varc = torch.where(vc[pick] > 0, varc, torch.zeros_like(varc))
        vk_all = torch.cat([vk, var])
        vc_all = torch.cat([vc, varc])
    else:
        vk_all, vc_all = vk, vc
    uk, inv = torch.unique(vk_all, return_inverse=True)
    uc = torch.zeros(uk.numel(), dtype=torch.int64, device=device)
    uc.index_add_(0, inv, vc_all)
    if cfg.high_count_kmers > 0 and uk.numel() > cfg.high_count_kmers:
        hidx = torch.randint(0, uk.numel(), (cfg.high_count_kmers,), generator=g, device=device)
        uc[hidx] = 100000 + hidx % 1000
    keep = uc >= 1
    uk, uc = uk[keep], uc[keep]
    perm = torch.randperm(uk.numel(), generator=g, device=device)
    keys, counts = uk[perm], uc[perm]
    # junction k-mers: windows straddling b1 / b2 (K-1 windows each)
    b1, b2 = tr["b1"][tid], tr["b2"][tid]
    strad = valid & (((b1 >= 0) & (pos < b1) & (pos + K > b1)) | ((b2 >= 0) & (pos < b2) & (pos + K > b2)))
    jk = torch.unique(allk[strad])
    # their SR counts (capped at 9999); absent / zero-count ones are dropped
    suk, sorder = torch.sort(keys)
    where = torch.searchsorted(suk, jk).clamp(max=suk.numel() - 1)
    found = suk[where] == jk
    jc = torch.where(found, counts[sorder[where]], torch.zeros_like(jk)).clamp(max=9999)
    jkeep = jc >= 1
    jk, jc = jk[jkeep], jc[jkeep]
    jperm = torch.randperm(jk.numel(), generator=g, device=device)
    return keys, counts, jk[jperm], jc[jperm]


def make_reads(cfg: SynthConfig, tr: dict, n_reads: int, device="cpu", seed_offset: int = 3, chunk: int = 100000):
    """ONT-like long reads: full-length transcripts sampled by coverage, errors 4:3:3, ASCII output."""
    g = _gen(cfg.seed * 1000 + seed_offset, device)
    tb, toff, cov, tlen = tr["t_bases"], tr["t_off"], tr["t_cov"], tr["tlen"]
    es, ei = cfg.read_error * 0.4, cfg.read_error * 0.3
    lut = torch.tensor(list(b"ACGT"), dtype=torch.uint8, device=device)
    outs, lens = [], []
    done = 0
    w = (cov / cov.sum()).to(torch.float64)
    while done < n_reads:
        n = min(chunk, n_reads - done)
        t = torch.multinomial(w, n, replacement=True, generator=g)
        L = tlen[t]
        roff = torch.zeros(n + 1, dtype=torch.int64, device=device)
        roff[1:] = torch.cumsum(L, 0)
        M = int(roff[-1])
        rid = torch.repeat_interleave(torch.arange(n, device=device), L)
        p = torch.arange(M, device=device) - roff[rid]
        base = tb[toff[t][rid] + p]
        u = torch.rand(M, generator=g, device=device)
        rb = torch.randint(0, 4, (M,), generator=g, device=device, dtype=torch.uint8)
        rd = torch.randint(1, 4, (M,), generator=g, device=device, dtype=torch.uint8)
        is_sub = u < es
        is_ins = (~is_sub) & (u < es + ei)
        is_del = (~is_sub) & (~is_ins) & (u < cfg.read_error)
        cnt = torch.ones(M, dtype=torch.int64, device=device)
        cnt[is_ins] = 2
        cnt[is_del] = 0
        ooff = torch.cumsum(cnt, 0) - cnt
        total = int(ooff[-1] + cnt[-1]) if M > 0 else 0
        out = torch.zeros(total, dtype=torch.uint8, device=device)
        b0 = torch.where(is_sub, (base + rd) & 3, base)
        keep = ~is_del
        out[ooff[keep]] = b0[keep]
        out[ooff[is_ins] + 1] = rb[is_ins]
        # read boundaries in output coordinates
        ends = torch.zeros(n + 1, dtype=torch.int64, device=device)
        first = roff[:-1]
        nz = L > 0
        ends[:-1][nz] = ooff[first[nz]]
        ends[-1] = total
        # empty transcripts cannot occur (len_min >= 300)
        outs.append(lut[out.to(torch.int64)])
        lens.append(ends[1:] - ends[:-1])
        done += n
    reads = torch.cat(outs)
    ln = torch.cat(lens)
    read_off = torch.zeros(n_reads + 1, dtype=torch.int64, device=device)
    read_off[1:] = torch.cumsum(ln, 0)
    return reads, read_off


def make_workload(cfg: SynthConfig, device="cpu", with_reads: bool = True) -> Workload:
    tr = make_transcriptome(cfg, device)
    keys, counts, jk, jc = make_counts(cfg, tr, device)
    w = Workload(cfg=cfg, t_bases=tr["t_bases"], t_off=tr["t_off"], t_cov=tr["t_cov"], keys=keys, counts=counts,
                 jkeys=jk, jcounts=jc)
    w.extra["tr"] = tr
    if with_reads:
        w.reads, w.read_off = make_reads(cfg, tr, cfg.n_reads, device)
    return w


# ---------------------------------------------------------------------------- text files for the CLI
def write_dump(path: str, keys: torch.Tensor, counts: torch.Tensor, k: int) -> None:
    """Jellyfish `dump -c` format: KMER<space>COUNT\\n, one line per k-mer, in the given order."""
    keys = keys.cpu().numpy().astype(np.int64)
    counts = counts.cpu().numpy().astype(np.int64)
    shifts = (2 * (k - 1 - np.arange(k))).astype(np.int64)
    with open(path, "wb") as f:
        step = 1 << 18
        for s in range(0, keys.shape[0], step):
            kk = keys[s:s + step]
            letters = BASES[(kk[:, None] >> shifts[None, :]) & 3]     # [m, k] uint8
            rows = letters.view(f"S{k}").reshape(-1)
            cc = counts[s:s + step]
            f.write(b"".join(r + b" " + str(c).encode() + b"\n" for r, c in zip(rows.tolist(), cc.tolist())))


def write_fasta(path: str, reads: torch.Tensor, read_off: torch.Tensor, prefix: str = "read_", width: int = 0) -> None:
    r = reads.cpu().numpy().tobytes()
    off = read_off.cpu().numpy()
    with open(path, "wb") as f:
        for i in range(len(off) - 1):
            f.write(b">" + prefix.encode() + str(i).encode() + b"\n")
            s = r[off[i]:off[i + 1]]
            if width and width > 0:
                for j in range(0, len(s), width):
                    f.write(s[j:j + width] + b"\n")
            else:
                f.write(s + b"\n")
