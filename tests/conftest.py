import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "hostemu"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def have_gpu() -> bool:
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


class Case:
    """A seeded synthetic workload plus the oracle's answer for it."""

    def __init__(self, cfg_index: int, scale: float, n_reads: int, junctions: bool = False, threads: int = 8, **over):
        from oracle import pyoracle as po
        from talc_b200 import synth
        cfg = synth.baseline_config(cfg_index, scale)
        cfg.n_reads = n_reads
        for k, v in over.items():
            setattr(cfg, k, v)
        self.cfg = cfg
        self.w = synth.make_workload(cfg)
        self.keys = self.w.keys.numpy().astype(np.uint64)
        self.counts = self.w.counts.numpy().astype(np.int64)
        self.jkeys = self.w.jkeys.numpy().astype(np.uint64) if junctions else None
        self.jcounts = self.w.jcounts.numpy().astype(np.int64) if junctions else None
        self.reads = self.w.reads.numpy()
        self.off = self.w.read_off.numpy().astype(np.uint64)
        self.op = po.make_params(k=cfg.k)
        self.otable = po.OracleTable(self.op).build_packed(self.keys, self.counts, self.jkeys, self.jcounts)
        self.o_out, self.o_off, self.o_status, self.o_ctr, _ = self.otable.correct(self.reads, self.off, threads=threads)

    def read(self, r: int) -> bytes:
        return self.reads[int(self.off[r]):int(self.off[r + 1])].tobytes()

    def oracle_read(self, r: int) -> bytes:
        return self.o_out[int(self.o_off[r]):int(self.o_off[r + 1])].tobytes()


_cases = {}


def get_case(name: str) -> Case:
    if name not in _cases:
        if name == "c1":
            _cases[name] = Case(1, 0.05, 250)
        elif name == "c3":
            _cases[name] = Case(3, 0.004, 250, junctions=True)
        elif name == "c5":
            _cases[name] = Case(5, 0.5, 160)
        elif name == "c1_cycle1":
            c = Case.__new__(Case)
            _cases[name] = c
            raise KeyError(name)
        else:
            raise KeyError(name)
    return _cases[name]


@pytest.fixture(scope="session")
def case_c1():
    return get_case("c1")


@pytest.fixture(scope="session")
def case_c3():
    return get_case("c3")


@pytest.fixture(scope="session")
def case_c5():
    return get_case("c5")
