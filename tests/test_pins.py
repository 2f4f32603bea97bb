"""The committed digests of the oracle's answers at the stated configuration sizes (tests/golden/oracle_sha256.json,
made by tools/make_pins.py): the oracle must reproduce them on this machine (drift guard), the array digest and the
`.fa` / `.log` file digests must be consistent, and the product's device code (host emulation) must reproduce the
k=31 extension and the junction case -- so the k=31 path is covered without a GPU as well."""
import hashlib
import json
import os
import sys

import numpy as np
import pytest

from conftest import ROOT, Case

sys.path.insert(0, os.path.join(ROOT, "tools"))
PINS = json.load(open(os.path.join(ROOT, "tests", "golden", "oracle_sha256.json")))


def _input_digest(case):
    h = hashlib.sha256()
    for a in (case.keys, case.counts, case.reads, case.off) + ((case.jkeys, case.jcounts) if case.jkeys is not None else ()):
        h.update(np.ascontiguousarray(a).tobytes())
    return h.hexdigest()


def _case(name):
    import make_pins
    ci, scale, n, usej, over = make_pins.PINS[name]
    return Case(ci, scale, n, junctions=usej, **over)


@pytest.mark.parametrize("name", ["config1", "config3_small", "k31"])
def test_oracle_reproduces_the_committed_digest(name):
    import make_pins
    pin = PINS[name]
    case = _case(name)
    if _input_digest(case) != pin["input_sha256"]:
        pytest.skip("torch generated a different workload on this machine; the pin does not apply")
    assert make_pins.array_digest(case.o_out, case.o_off, case.o_status) == pin["sha256"]
    ids = [b"read_%d" % r for r in range(len(case.off) - 1)]
    assert hashlib.sha256(make_pins.fasta_bytes(case.o_out, case.o_off, ids)).hexdigest() == pin["fa_sha256"]
    assert hashlib.sha256(b"".join(sorted(make_pins.log_lines(case.o_status, ids)))).hexdigest() == pin["log_sha256"]
    assert case.otable.size() == pin["table_entries_kept"]


@pytest.mark.parametrize("name", ["k31", "config3_small"])
def test_device_code_on_host_reproduces_the_pinned_cases(name):
    """The TALC_HD device code compiled with g++ (tests/hostemu): k=31 uses 62 key bits -- the whole u64 k-mer
    arithmetic (masks, successor shifts, reverse k-mers of the LEFT-ward cycle test) at its limit."""
    import make_pins
    import pyemu
    case = _case(name)
    et = pyemu.EmuTable(pyemu.params_from(case.op), case.keys, case.counts, case.jkeys, case.jcounts)
    out, off, st, ctr = et.correct(case.reads, case.off, arena_bytes=4 << 20, wide=True)
    assert np.array_equal(st, case.o_status)
    assert np.array_equal(off, case.o_off) and np.array_equal(out, case.o_out)
    for k in ("lookups_walk", "steps_inner", "steps_border", "cells_nw", "cells_xdrop", "gaps_bridged", "ev_gardening"):
        assert ctr[k] == case.o_ctr[k], k
    if _input_digest(case) == PINS[name]["input_sha256"]:
        assert make_pins.array_digest(out, off, st) == PINS[name]["sha256"]


def test_pin_script_runs_end_to_end(tmp_path):
    """tools/pin_reference.py is the one command that pins the oracle the day a reference binary exists.  Its mechanics
    (text inputs written, both programs run on them, .fa / sorted .log / .config.txt and the committed digests
    compared, exit code) are exercised here with the oracle's own front end standing in for the reference binary --
    this proves the script, not the oracle."""
    import subprocess
    import sys
    from oracle import pyoracle as po
    po.build()
    root = ROOT
    pr = subprocess.run([sys.executable, os.path.join(root, "tools", "pin_reference.py"), "--ref", po.BIN, "--configs", "3small",
                         "--workdir", str(tmp_path)], capture_output=True, text=True, cwd=root)
    assert pr.returncode == 0, pr.stdout + pr.stderr
    assert "PINNED" in pr.stdout and "DIFFERENT" not in pr.stdout
    # and it says so when the two differ: a "reference" that drops the last read
    fake = tmp_path / "fake_talc.sh"
    fake.write_text("#!/bin/sh\n%s \"$@\" || exit $?\nfor a in \"$@\"; do if [ \"$prev\" = -o ]; then o=$a; fi; prev=$a; done\n"
                    "head -n -2 $o.fa > $o.fa.tmp && mv $o.fa.tmp $o.fa\n" % po.BIN)
    fake.chmod(0o755)
    pr = subprocess.run([sys.executable, os.path.join(root, "tools", "pin_reference.py"), "--ref", str(fake), "--configs", "3small",
                         "--workdir", str(tmp_path)], capture_output=True, text=True, cwd=root)
    assert pr.returncode == 1 and "NOT PINNED" in pr.stdout
