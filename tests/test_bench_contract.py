"""bench.py contract pieces that can be checked without a GPU."""
import importlib.util
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _load_bench():
    spec = importlib.util.spec_from_file_location("bench_module", os.path.join(ROOT, "bench.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_reference_arm_reports_unavailable_without_cuda():
    import torch

    if torch.cuda.is_available():
        return  # on a GPU box the arm really runs; the driver exercises it
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference"], capture_output=True,
                         text=True, timeout=600)
    assert out.returncode == 0
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and "unavailable" in line


def test_reference_probe_decision(monkeypatch):
    bench = _load_bench()

    class Proc:
        def __init__(self, stdout):
            self.stdout, self.stderr, self.returncode = stdout, "", 0

    monkeypatch.delenv("DISABLE_MMA_V5", raising=False)
    tmem = '{"probe_error": "OutOfResources: out of resource: tensor memory, Required: 704, Hardware limit: 512"}\n'
    monkeypatch.setattr(bench.subprocess, "run", lambda *a, **k: Proc("log line\n" + tmem))
    assert "DISABLE_MMA_V5" in bench.probe_reference_backward(0)
    monkeypatch.setattr(bench.subprocess, "run", lambda *a, **k: Proc('{"probe_ok": true}\n'))
    assert bench.probe_reference_backward(0) == {}
    monkeypatch.setattr(bench.subprocess, "run", lambda *a, **k: Proc('{"probe_error": "ImportError: no triton"}\n'))
    assert bench.probe_reference_backward(0) == {}  # any other failure: leave the environment alone
    monkeypatch.setenv("DISABLE_MMA_V5", "1")
    assert bench.probe_reference_backward(0) == {}  # the user already chose
