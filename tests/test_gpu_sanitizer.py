"""compute-sanitizer memcheck over the smoke path: one golden train + eval pass in both FC modes
(FFMA kernels and the TMA / tcgen05 kernels) and three fused training steps (CUDA graph, side stream, row-lazy Adam).
Out-of-bounds or misaligned global / shared accesses of any kernel fail the test."""
import os
import shutil
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


@pytest.mark.skipif(not torch.cuda.is_available(), reason="needs a GPU")
def test_memcheck_smoke_path():
    exe = shutil.which("compute-sanitizer") or "/usr/local/cuda/bin/compute-sanitizer"
    if not os.path.exists(exe):
        pytest.skip("compute-sanitizer is not installed")
    cmd = [exe, "--tool", "memcheck", "--error-exitcode", "9", "--launch-timeout", "0", sys.executable,
           os.path.join(ROOT, "__graft_entry__.py"), "smoke"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
    tail = (r.stdout[-3000:] + r.stderr[-3000:])
    assert r.returncode == 0 and "ERROR SUMMARY: 0 errors" in (r.stdout + r.stderr) and "smoke ok" in r.stdout, tail
