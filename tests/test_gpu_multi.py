"""Multi-GPU paths on CUDA (skipped with fewer than two devices -- run with `gpurun --gpus 2`):
table replicas made by export -> import (the NCCL broadcast path of bench.py), by talc_table_copy (peer copy) and by
the library's own NCCL broadcast answer look-ups identically and correct to the same bytes; `talc --gpus 2` writes the
same files as `talc --gpus 1`."""
import os
import subprocess

import numpy as np
import pytest

from conftest import have_gpu

pytestmark = pytest.mark.gpu


def _ndev():
    import torch
    return torch.cuda.device_count() if have_gpu() else 0


@pytest.fixture(scope="module")
def api():
    if not have_gpu():
        pytest.fail("GPU tests selected but no CUDA device is visible")
    if _ndev() < 2:
        pytest.skip("needs two CUDA devices")
    from talc_b200 import api as _api
    return _api


def _same_answers(case, a, b):
    rng = np.random.default_rng(1)
    probe = np.concatenate([case.keys[rng.integers(0, len(case.keys), 9000)], rng.integers(0, 1 << 42, 1000).astype(np.uint64)])
    ca, la, fa = a.lookup(probe)
    cb, lb, fb = b.lookup(probe)
    assert np.array_equal(ca, cb) and np.array_equal(la, lb) and np.array_equal(fa, fb)
    assert (la > 0).any()
    out, off, st, ctr = b.correct(case.reads, case.off)
    assert np.array_equal(st, case.o_status)
    assert np.array_equal(off, case.o_off) and np.array_equal(out, case.o_out)


def test_export_import_replica(api, case_c3):
    import torch
    case = case_c3
    a = api.Talc(api.default_params(case.cfg.k), device=0)
    a.load_packed(case.keys, case.counts, case.jkeys, case.jcounts)
    info = a.table_info()
    staging0 = torch.empty(info["capacity"] * 16, dtype=torch.uint8, device="cuda:0")
    a.table_export_device(staging0)
    staging1 = staging0.to("cuda:1")  # stands in for the NCCL broadcast of bench.py
    b = api.Talc(api.default_params(case.cfg.k), device=1)
    b.table_import_device(staging1, info["capacity"], info["entries"])
    assert b.table_info()["entries"] == info["entries"]
    _same_answers(case, a, b)


def test_peer_copy_replica(api, case_c3):
    case = case_c3
    a = api.Talc(api.default_params(case.cfg.k), device=0)
    a.load_packed(case.keys, case.counts, case.jkeys, case.jcounts)
    b = api.Talc(api.default_params(case.cfg.k), device=1)
    api.table_copy(b, a)
    _same_answers(case, a, b)


def test_library_nccl_replicate(api, case_c3):
    """talc_table_replicate: ONE ncclBroadcast of the slot array issued by the library itself."""
    case = case_c3
    a = api.Talc(api.default_params(case.cfg.k), device=0)
    a.load_packed(case.keys, case.counts, case.jkeys, case.jcounts)
    b = api.Talc(api.default_params(case.cfg.k), device=1)
    ms = api.table_replicate([a, b])
    assert ms >= 0
    _same_answers(case, a, b)


def test_cli_two_gpus_equals_one(api, tmp_path):
    import torch
    from talc_b200 import build, synth
    cfg = synth.baseline_config(1, 0.05)
    cfg.n_reads = 600
    w = synth.make_workload(cfg)
    synth.write_dump(str(tmp_path / "sr.dump"), w.keys, w.counts, cfg.k)
    synth.write_fasta(str(tmp_path / "reads.fa"), w.reads, w.read_off, width=80)
    cli = build.build_cli()
    args = [str(tmp_path / "reads.fa"), "--SRCounts", str(tmp_path / "sr.dump"), "-k", str(cfg.k)]
    assert subprocess.call([cli] + args + ["-o", "one", "--gpus", "1"], cwd=tmp_path, stdout=subprocess.DEVNULL) == 0
    assert subprocess.call([cli] + args + ["-o", "two", "--gpus", "2", "--batch-reads", "100"], cwd=tmp_path,
                           stdout=subprocess.DEVNULL) == 0
    assert subprocess.call([cli] + args + ["-o", "peer", "--gpus", "2", "--batch-reads", "77", "--replicate", "peer"], cwd=tmp_path,
                           stdout=subprocess.DEVNULL) == 0
    for ext in (".fa", ".log"):
        a = open(tmp_path / ("one" + ext), "rb").read() if os.path.exists(tmp_path / ("one" + ext)) else b""
        for other in ("two", "peer"):  # NCCL broadcast, then peer copies
            b = open(tmp_path / (other + ext), "rb").read() if os.path.exists(tmp_path / (other + ext)) else b""
            assert a == b, (other, ext)
    assert len(open(tmp_path / "one.fa", "rb").read()) > int(w.read_off[-1])
