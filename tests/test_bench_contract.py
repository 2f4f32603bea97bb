"""bench.py without a GPU: the pieces of its output contract that are host logic -- the committed ncu capture it
reads for `roofline.traffic` / `roofline.issue`, the peak it divides by, exactly one JSON line on stdout whatever
the libraries underneath print, and a loud failure (no numbers) when there is no CUDA device."""
import importlib.util
import json
import os
import subprocess
import sys

from conftest import ROOT


def _bench():
    spec = importlib.util.spec_from_file_location("talc_bench", os.path.join(ROOT, "bench.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def test_committed_capture_is_readable():
    b = _bench()
    t, n = b.measured_traffic(), b.measured_instructions()
    assert t is not None and 30e9 < t < 80e9            # DRAM bytes of one correct_kernel launch of the bench workload
    assert n is not None and 1e11 < n < 5e11            # its warp-instructions
    assert os.path.exists(os.path.join(ROOT, b.TRAFFIC_SOURCE.split(" ")[0]))
    peak, src = b.peaks()
    assert 3000 < peak < 9000 and src


def test_one_json_line_on_stdout_whatever_else_is_printed(tmp_path):
    """_claim_stdout keeps a private duplicate of fd 1 for the JSON line and points fd 1 at stderr: a C library that
    prints to stdout (NCCL's version banner) cannot add lines."""
    code = ("import os, sys, ctypes; sys.path.insert(0, %r); import importlib.util as u;"
            "s = u.spec_from_file_location('b', %r); b = u.module_from_spec(s); s.loader.exec_module(b);"
            "b._claim_stdout(); print('python noise'); ctypes.CDLL(None).puts(b'C stdio noise'); ctypes.CDLL(None).fflush(None);"
            "os.write(1, b'raw fd noise\\n'); b.emit({'metric': 'm', 'value': 1.5})") % (ROOT, os.path.join(ROOT, "bench.py"))
    pr = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True)
    assert pr.returncode == 0, pr.stderr
    lines = pr.stdout.splitlines()
    assert len(lines) == 1 and json.loads(lines[0]) == {"metric": "m", "value": 1.5}
    assert "python noise" in pr.stderr and "C stdio noise" in pr.stderr and "raw fd noise" in pr.stderr


def test_no_numbers_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        import pytest
        pytest.skip("needs a machine without a CUDA device")
    pr = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--warmup", "1"], capture_output=True,
                        text=True, cwd=ROOT)
    assert pr.returncode != 0 and pr.stdout.strip() == "" and "no CPU fallback" in pr.stderr
